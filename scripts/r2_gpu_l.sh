#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2l_gpus.txt
timeout 600 python -m pytest tests/test_dispatch_gpu.py -q -m gpu -rA 2>&1 | tail -12 > gpurun_out/r2l_dispatch_test.log
cat gpurun_out/r2l_dispatch_test.log
timeout 900 python bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_2gpu.json').read().strip().splitlines()[-1])
print('2 GPUs:', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))
for k,v in d.get('secondary',{}).items():
    print('   ', k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','unit','ms_per_step','error','videos_per_sec','n_gpus')})
PY
tail -5 gpurun_out/r2l_bench_2gpu.err
