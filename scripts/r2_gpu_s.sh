#!/bin/bash
# GPU call S: asynchronous host calls -- parity tests, headline bench (e2e now pipelined), c5 at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_clip_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2s_tests.txt
cat gpurun_out/r2s_tests.txt
python bench.py --no-cpu --no-secondary > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['e2e'].get('ms_per_step'))"
VF_C5_TRACE=1 python bench.py --workload c5 --no-cpu > gpurun_out/r2s_c5_10k.json 2> gpurun_out/r2s_c5_10k.err
grep "c5 trace" gpurun_out/r2s_c5_10k.err; python -c "
import json; d=json.loads(open('gpurun_out/r2s_c5_10k.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
VF_BENCH_C5_VIDEOS=1250 VF_C5_TRACE=1 python bench.py --workload c5 --no-cpu > gpurun_out/r2s_c5_1250.json 2> gpurun_out/r2s_c5_1250.err
grep "c5 trace" gpurun_out/r2s_c5_1250.err; python -c "
import json; d=json.loads(open('gpurun_out/r2s_c5_1250.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
