#!/bin/bash
# I3D + RAFT parity tests on the current build
cd "$(dirname "$0")/.."
timeout -s KILL 600 python -m pytest tests/test_raft_gpu.py tests/test_i3d_gpu.py tests/test_extract_i3d_raft_gpu.py -q -m gpu 2>&1 | grep -v -i warn | tail -6
