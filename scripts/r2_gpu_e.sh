#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_trace.so timeout 300 python scripts/gemm_trace.py > gpurun_out/r2e_gemm_trace.txt 2>&1
echo "######## no epilogue (accumulators released unread)" >> gpurun_out/r2e_gemm_trace.txt
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_tracenoepi.so timeout 300 python scripts/gemm_trace.py >> gpurun_out/r2e_gemm_trace.txt 2>&1
for c in 125 200 334 500; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary --chunk $c > gpurun_out/r2e_bench_chunk$c.json 2> gpurun_out/r2e_bench_chunk$c.err
done
cat gpurun_out/r2e_gemm_trace.txt
