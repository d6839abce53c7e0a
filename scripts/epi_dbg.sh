#!/bin/bash
# development probe: epilogue-only GEMM timings under the VF_GEMM_DBG switches
for mode in pair 1cta; do for d in 0 1 2 3 4; do
  echo "== $mode dbg=$d"; VF_GEMM=$mode VF_GEMM_DBG=$d timeout -s KILL 100 python scripts/epi_sweep.py 2>&1 | grep "K=  64"
done; done
