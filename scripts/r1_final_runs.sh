#!/bin/bash
# Round-1 measurement pass on the GPU box: tests, the three bench workloads, the reference arm, ncu captures.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu 2>&1 | grep -v -i warn | tail -5 > gpurun_out/r1_pytest_gpu.txt
timeout -s KILL 400 python bench.py --steps 100 --warmup 5 > gpurun_out/r1_bench_clip.json 2> gpurun_out/r1_bench_clip.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_clip_reference.json 2>> gpurun_out/r1_bench_clip.err
timeout -s KILL 300 python bench.py --workload i3d --steps 10 --warmup 3 > gpurun_out/r1_bench_i3d.json 2> gpurun_out/r1_bench_i3d.err
timeout -s KILL 400 python bench.py --workload raft --steps 5 --warmup 3 > gpurun_out/r1_bench_raft.json 2> gpurun_out/r1_bench_raft.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1092 -c 728 --csv --log-file gpurun_out/r1_launches_clip.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r1_ncu_bench.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_pair -s 60 -c 4 -o gpurun_out/r1_prof_gemm \
    python scripts/ncu_clip_once.py 250 250 > gpurun_out/r1_ncu_full.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size \
    --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches_i3d.csv env VF_BENCH_I3D_STACKS=8 python bench.py --workload i3d --steps 1 --warmup 1 --no-cpu > gpurun_out/r1_ncu_i3d.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_raft.csv \
    python scripts/ncu_raft_once.py 9 > gpurun_out/r1_ncu_raft.log 2>&1
cat gpurun_out/r1_pytest_gpu.txt
python - <<'PY'
import json
for f in ("r1_bench_clip", "r1_bench_clip_reference", "r1_bench_i3d", "r1_bench_raft"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").readline())
        print(f, d.get("value"), d.get("unit"), "e2e", d.get("e2e", {}).get("value"), "roof", d.get("roofline", {}).get("frac"),
              "cpu", d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
