#!/bin/bash
# GPU call R: host timeline of the list path (c5) at N=1, 10k and 1250 videos
mkdir -p gpurun_out
VF_C5_TRACE=1 python bench.py --workload c5 --no-cpu > gpurun_out/r2r_c5_10k.json 2> gpurun_out/r2r_c5_10k.err
grep "c5 trace" gpurun_out/r2r_c5_10k.err; python -c "
import json; d=json.loads(open('gpurun_out/r2r_c5_10k.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
VF_BENCH_C5_VIDEOS=1250 VF_C5_TRACE=1 python bench.py --workload c5 --no-cpu > gpurun_out/r2r_c5_1250.json 2> gpurun_out/r2r_c5_1250.err
grep "c5 trace" gpurun_out/r2r_c5_1250.err; python -c "
import json; d=json.loads(open('gpurun_out/r2r_c5_1250.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
