#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_i3d_gpu.py tests/test_extract_i3d_raft_gpu.py -q -m gpu -s 2>&1 | grep -E "rel-L2|rel |passed|failed|Error|assert" | tee gpurun_out/r2n_i3d_tests.log
for pol in default none; do
  VF_I3D_SINGLE=$pol timeout 400 python bench.py --workload i3d --steps 10 --warmup 3 --no-cpu > gpurun_out/r2n_bench_i3d_$pol.json 2> gpurun_out/r2n_bench_i3d_$pol.err
done
python - <<'PY'
import json
for f in ('r2n_bench_i3d_default','r2n_bench_i3d_none'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'exec/alg', round(r['executed_over_algorithmic'],2))
PY
