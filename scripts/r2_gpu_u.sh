#!/bin/bash
# GPU call U: list path (c5) stage waits + device timeline at N=1
mkdir -p gpurun_out
for cfg in "8 1000 1" "8 1000 0"; do
  set -- $cfg
  VF_DECODE_WORKERS=$1 VF_CLIP_BATCH_FRAMES=$2 VF_C5_TRACE=$3 python bench.py --workload c5 --no-cpu > gpurun_out/r2u_c5_w$1_b$2_t$3.json 2> gpurun_out/r2u_c5_w$1_b$2_t$3.err
  echo "workers $1 batch $2 trace $3:"; grep "c5 trace\|AssertionError\|Error" gpurun_out/r2u_c5_w$1_b$2_t$3.err | cut -c1-420
  python -c "
import json; d=json.loads(open('gpurun_out/r2u_c5_w$1_b$2_t$3.json').read().strip().splitlines()[-1]); print('   ', round(d['value']), d['ms_per_step'])"
done
exit 0
