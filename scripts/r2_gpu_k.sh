#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2k_tests.log 2>&1
tail -3 gpurun_out/r2k_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for tag in ln5 ln6; do
  L=$PWD/video_features_b200/libvfeat.so; [ $tag = ln6 ] && L=$PWD/video_features_b200/libvfeat_ln6.so
  VF_LIBVFEAT=$L timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2k_bench_$tag.json 2> gpurun_out/r2k_bench_$tag.err
done
for cfg in "8 1024" "12 1024" "16 1024" "8 2048" "16 2048"; do
  set -- $cfg
  VF_DECODE_WORKERS=$1 VF_CLIP_BATCH_FRAMES=$2 timeout 300 python bench.py --workload c5 > gpurun_out/r2k_c5_w$1_b$2.json 2> gpurun_out/r2k_c5_w$1_b$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), {k:round(v,3) for k,v in r['eager_ms_per_step_by_kernel'].items()})
for f in sorted(glob.glob('gpurun_out/r2k_c5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'frames/s', round(d['videos_per_sec']), 'videos/s wall', round(d['host_wall_s_rank0'],2))
    except Exception as e: print(f, 'ERR', e)
PY
