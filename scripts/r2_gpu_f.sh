#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2f_gemm_trace.txt
for lv in 0 1 2 3 4 9; do
  echo "######## VF_DBG_EPI=$lv (0 all, 1 no TMA store, 2 TMEM load + math, 3 TMEM load only, 4 all but the TMEM load, 9 none)" >> gpurun_out/r2f_gemm_trace.txt
  VF_LIBVFEAT=$PWD/video_features_b200/libvfeat_tr$lv.so timeout 300 python scripts/gemm_trace.py 2>&1 | grep -E "^==|^      [2345] |entry ->" >> gpurun_out/r2f_gemm_trace.txt
done
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_clip_gpu.py tests/test_i3d_gpu.py tests/test_raft_gpu.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log
cat gpurun_out/r2f_gemm_trace.txt
