#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_clip_gpu.py -k fused -x -q -s > gpurun_out/r2j_fused_test.log 2>&1
rc=$?
echo "fused unit test rc=$rc" >> gpurun_out/r2j_fused_test.log
tail -22 gpurun_out/r2j_fused_test.log
if [ $rc -ne 0 ]; then exit 0; fi
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_attntr.so timeout 200 python scripts/attn_trace.py 2>&1 | tee gpurun_out/r2j_attn_trace.txt
timeout 600 python -m pytest tests/test_clip_gpu.py tests/test_extract_clip_gpu.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2j_tests.log
for mode in fused split; do
  VF_CLIP_ATTN=$mode timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2j_bench_$mode.json 2> gpurun_out/r2j_bench_$mode.err
done
python - <<'PY'
import json
for f in ('r2j_bench_fused','r2j_bench_split'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), {k:round(v,3) for k,v in r['eager_ms_per_step_by_kernel'].items()})
PY
