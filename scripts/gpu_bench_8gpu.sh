#!/bin/bash
# 8-GPU run of the default bench line (torchrun, one rank per GPU), as the driver launches it for its scaling curve
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_8gpu_ngpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "rc=$?" >> gpurun_out/r2_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu.json').read().strip().splitlines()[-1])
print('8 GPUs:', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), d['clocks'])
for k,v in d.get('secondary',{}).items():
    print('   ', k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','unit','ms_per_step','error','videos_per_sec','n_gpus','host_wall_s_rank0')})
PY
grep "c5 trace rank [07]\]" gpurun_out/r2_bench_8gpu.err | cut -c1-400; tail -2 gpurun_out/r2_bench_8gpu.err
