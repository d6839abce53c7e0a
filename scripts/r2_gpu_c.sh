#!/bin/bash
# Round 2, GPU call C: register-direct epilogue stores (mode 6), deeper TMA ring; full -m gpu suite; default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c_gemm_variants.txt
: > $OUT
for tag in m5g4 m5g4s4 m6g2 m6g3 m6g4 m6g4s6; do
  echo "=== $tag" >> $OUT
  VF_GEMM_EPI3=$tag VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python scripts/gemm_shapes.py >> $OUT 2>&1
done
VF_GEMM=m6g4 VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_m6g4.so timeout 300 python scripts/gemm_sweep.py >> $OUT 2>&1
: > gpurun_out/r2c_tests.log
for tag in m6g4 m6g4s6; do
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_clip_gpu.py tests/test_i3d_gpu.py tests/test_raft_gpu.py -q -m gpu 2>&1 | tail -8 >> gpurun_out/r2c_tests.log
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_$tag.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.err
done
timeout 1500 python -m pytest tests -q -m gpu -rA --durations=10 > gpurun_out/r2c_tests_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_tests_full.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -3 gpurun_out/r2c_tests_full.log; cat gpurun_out/r2c_tests.log
