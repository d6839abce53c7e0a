#!/bin/bash
# Last check of a round: smoke(), the whole -m gpu suite and the list-path line on the committed build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v -i warn | tail -6
timeout -s KILL 900 python -m pytest tests -q -m gpu 2>&1 | grep -v -i warn | tail -4 | tee gpurun_out/r2_pytest_gpu.txt
timeout -s KILL 300 python bench.py --workload c5 --no-cpu > gpurun_out/r2_c5_final.json 2> gpurun_out/r2_c5_final.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c5_final.json').read().strip().splitlines()[-1]); print('c5', round(d['value']), d['ms_per_step'])"
