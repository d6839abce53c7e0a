#!/bin/bash
# GPU call P: ViT-B/16 tower -- parity tests, throughput by chunk size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_clip_gpu.py -x -q -m gpu -k b16 2>&1 | tail -15 > gpurun_out/r2p_tests.txt
cat gpurun_out/r2p_tests.txt
timeout 600 python scripts/b16_time.py 1008 63,126 2>&1 | tee gpurun_out/r2p_b16_time.txt | tail -8
