#!/bin/bash
# Round-2 measurement pass on the GPU box: tests, the default bench line (with its secondary workloads), the reference
# arm, the torch-on-GPU library bar, ncu launch list.  Outputs -> gpurun_out/r2_*; scripts/make_profile_summary_r2.py
# turns them into profiles/r2_summary.md.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v -i warn | tail -4 > gpurun_out/r2_pytest_gpu.txt
timeout -s KILL 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench_clip.json 2> gpurun_out/r2_bench_clip.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_clip_reference.json 2>> gpurun_out/r2_bench_clip.err
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --torch-gpu > gpurun_out/r2_bench_clip_torchgpu.json 2>> gpurun_out/r2_bench_clip.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_clip.csv \
    python scripts/ncu_clip_once.py 1000 250 > gpurun_out/r2_ncu_list.log 2>&1
timeout -s KILL 300 python scripts/b16_time.py 1008 126 > gpurun_out/r2_b16_time.txt 2>&1
# is the NVDEC user-mode library on the box at all?  headers / a demuxer are not in the image either way
(ldconfig -p | grep -i -E "nvcuvid|nvidia-encode" || echo "libnvcuvid: not found by ldconfig"; ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvcuvid|nvidia-encode" || true) > gpurun_out/r2_nvdec_probe.txt 2>&1
cat gpurun_out/r2_pytest_gpu.txt gpurun_out/r2_b16_time.txt gpurun_out/r2_nvdec_probe.txt
python - <<'PY'
import json
for f in ("r2_bench_clip", "r2_bench_clip_reference", "r2_bench_clip_torchgpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d.get("value"), 1), d.get("unit"), "e2e", round(d.get("e2e", {}).get("value", 0), 1), "roof", d.get("roofline", {}).get("frac"),
              "cpu", d.get("cpu_baseline", {}).get("value"), d.get("torch_gpu_baseline"))
        for k, v in d.get("secondary", {}).items():
            print("   ", k, v.get("value"), v.get("unit"), v.get("error"))
    except Exception as e:
        print(f, "FAILED", e)
PY
