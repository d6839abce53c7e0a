#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_trace.so timeout 300 python scripts/gemm_trace.py > gpurun_out/r2d_gemm_trace.txt 2>&1
VF_GEMM_EPI3=main VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_main.so timeout 300 python scripts/gemm_shapes.py >> gpurun_out/r2d_gemm_trace.txt 2>&1
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_main.so timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2d_tests.log 2>&1
tail -3 gpurun_out/r2d_tests.log
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_main.so timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_gemm_trace.txt
