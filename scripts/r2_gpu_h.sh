#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_clip_gpu.py -k fused -x -q -s > gpurun_out/r2h_fused_test.log 2>&1
rc=$?
echo "fused unit test rc=$rc" >> gpurun_out/r2h_fused_test.log
tail -25 gpurun_out/r2h_fused_test.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_clip_gpu.py tests/test_extract_clip_gpu.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
for mode in fused split; do
  VF_CLIP_ATTN=$mode timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2h_bench_$mode.json 2> gpurun_out/r2h_bench_$mode.err
done
for resid in y mix; do
  VF_CLIP_RESID=$resid timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2h_bench_resid_$resid.json 2> gpurun_out/r2h_bench_resid_$resid.err
done
