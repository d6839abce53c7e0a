#!/bin/bash
# bench lines with the torch-eager-on-GPU library baseline (key torch_gpu_baseline)
timeout -s KILL 200 python bench.py --steps 10 --warmup 5 --no-cpu --torch-gpu > gpurun_out/r1_torchgpu_clip.json 2> gpurun_out/r1_torchgpu.err
timeout -s KILL 200 python bench.py --workload i3d --steps 4 --warmup 3 --no-cpu --torch-gpu > gpurun_out/r1_torchgpu_i3d.json 2>> gpurun_out/r1_torchgpu.err
timeout -s KILL 300 python bench.py --workload raft --steps 2 --warmup 3 --no-cpu --torch-gpu > gpurun_out/r1_torchgpu_raft.json 2>> gpurun_out/r1_torchgpu.err
python - <<'PY'
import json
for w in ("clip", "i3d", "raft"):
    try:
        d = json.loads(open(f"gpurun_out/r1_torchgpu_{w}.json").readline())
        print(w, round(d["value"], 1), d["unit"], d.get("torch_gpu_baseline"))
    except Exception as e:
        print(w, "FAILED", e)
PY
tail -3 gpurun_out/r1_torchgpu.err
