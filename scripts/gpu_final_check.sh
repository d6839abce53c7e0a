#!/bin/bash
# Last check of a round: smoke() and the whole -m gpu suite on the committed build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v -i warn | tail -6
timeout -s KILL 900 python -m pytest tests -q -m gpu 2>&1 | grep -v -i warn | tail -4 | tee gpurun_out/r2_pytest_gpu.txt
