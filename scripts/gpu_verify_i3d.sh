#!/bin/bash
# I3D parity tests on the current build, I3D rgb / flow throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_i3d_gpu.py tests/test_extract_i3d_raft_gpu.py -q -m gpu 2>&1 | grep -v -i warn | tail -5
python scripts/ncu_i3d_once.py rgb 32 | tail -1
python scripts/ncu_i3d_once.py rgb 8 | tail -1
