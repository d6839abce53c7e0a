#!/bin/bash
# Round 2, GPU call B: epilogue warp-group count x hand-off variants; residual add in the epilogue (TMA reduction) vs fp16 y.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2b_gemm_variants.txt
: > $OUT
for tag in m0g2 m5g2 m0g3 m5g3 m0g4 m5g4; do
  echo "=== $tag" >> $OUT
  VF_GEMM_EPI3=$tag VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python scripts/gemm_shapes.py >> $OUT 2>&1
done
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_clip_gpu.py tests/test_transform_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2b_tests.log
for tag in m5g4 m0g4; do
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_clip_gpu.py -q -m gpu 2>&1 | tail -5 >> gpurun_out/r2b_tests.log
done
for resid in y acc; do
  VF_CLIP_RESID=$resid timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2b_bench_resid_$resid.json 2> gpurun_out/r2b_bench_resid_$resid.err
done
for tag in m5g4 m0g4 m5g3; do
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2b_bench_$tag.json 2> gpurun_out/r2b_bench_$tag.err
done
cat gpurun_out/r2b_tests.log
