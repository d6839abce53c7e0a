#!/bin/bash
# Full -m gpu suite on the current build + one ncu --set full capture of the ViT-B/16 attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu 2>&1 | grep -v -i warn | tail -4 > gpurun_out/r2_pytest_gpu.txt
cat gpurun_out/r2_pytest_gpu.txt
timeout -s KILL 300 ncu --set full --clock-control none -k regex:attention_long -s 12 -c 1 --csv --page raw --log-file gpurun_out/r2_prof_attn_long_raw.csv \
    python scripts/b16_time.py 252 126 > gpurun_out/r2_ncu_b16.log 2>&1
tail -2 gpurun_out/r2_ncu_b16.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2_prof_attn_long_raw.csv", errors="replace") if l.startswith('"'))]
if len(rows) >= 3:
    hdr, units, val = rows[0], rows[1], rows[2]
    for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "launch__registers_per_thread",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
        if k in hdr:
            print(k, val[hdr.index(k)], units[hdr.index(k)])
PY
