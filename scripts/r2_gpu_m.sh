#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VF_TAG=loaders4-7 timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | tee gpurun_out/r2m_attn_time.txt
VF_TAG=loaders2-5 VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_l2.so timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | tee -a gpurun_out/r2m_attn_time.txt
for tag in main l2; do
  L=$PWD/video_features_b200/libvfeat.so; [ $tag = l2 ] && L=$PWD/video_features_b200/libvfeat_l2.so
  VF_LIBVFEAT=$L timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2m_bench_$tag.json 2> gpurun_out/r2m_bench_$tag.err
done
python - <<'PY'
import json
for f in ('r2m_bench_main','r2m_bench_l2'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), {k:round(v,3) for k,v in r['eager_ms_per_step_by_kernel'].items()}, r.get('qkv_attention',{}).get('frac'))
PY
# ---- ncu: launch list of two steps, then full captures of the kernels of one 250-frame chunk
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_clip.csv \
    python scripts/ncu_clip_once.py 1000 250 > gpurun_out/r2m_ncu_list.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_pair -s 40 -c 4 -o gpurun_out/r2_prof_gemm \
    python scripts/ncu_clip_once.py 250 250 > gpurun_out/r2m_ncu_gemm.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:qkv_attention -s 6 -c 1 -o gpurun_out/r2_prof_attn \
    python scripts/ncu_clip_once.py 250 250 > gpurun_out/r2m_ncu_attn.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:add_layernorm -s 20 -c 1 -o gpurun_out/r2_prof_ln \
    python scripts/ncu_clip_once.py 250 250 > gpurun_out/r2m_ncu_ln.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:resample -s 2 -c 2 -o gpurun_out/r2_prof_resample \
    python scripts/ncu_resize_once.py > gpurun_out/r2m_ncu_resample.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
