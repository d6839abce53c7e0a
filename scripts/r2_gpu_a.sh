#!/bin/bash
# Round 2, GPU call A: store-path micro-benchmark, GEMM epilogue variants, the -m gpu suite, a first bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 ./scripts/ubench/store_paths > gpurun_out/r2a_store_paths.txt 2>&1
for tag in epi0 epi4 epi5 noepi nostore0 nostore5; do
  echo "=== $tag" >> gpurun_out/r2a_gemm_variants.txt
  VF_GEMM=$tag VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python scripts/gemm_sweep.py >> gpurun_out/r2a_gemm_variants.txt 2>&1
  VF_GEMM_EPI3=$tag VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python scripts/gemm_shapes.py >> gpurun_out/r2a_gemm_variants.txt 2>&1
done
for tag in epi4 epi5; do
  echo "=== tests with $tag" >> gpurun_out/r2a_gemm_variants.txt
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_clip_gpu.py -q -m gpu 2>&1 | tail -15 >> gpurun_out/r2a_gemm_variants.txt
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2a_bench_$tag.json 2> gpurun_out/r2a_bench_$tag.err
done
timeout 1500 python -m pytest tests -q -m gpu -rA --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.log
