#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_tr0.so timeout 300 python scripts/gemm_trace.py 2>&1 | grep -E "^==|^      [2345] |entry ->" > gpurun_out/r2g_gemm_trace.txt
VF_GEMM_EPI3=stg timeout 300 python scripts/gemm_shapes.py >> gpurun_out/r2g_gemm_trace.txt 2>&1
VF_GEMM=stg timeout 300 python scripts/gemm_sweep.py >> gpurun_out/r2g_gemm_trace.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2g_tests.log 2>&1
tail -4 gpurun_out/r2g_tests.log
VF_NO_PDL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2g_bench_nopdl.json 2> gpurun_out/r2g_bench_nopdl.err
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
cat gpurun_out/r2g_gemm_trace.txt
