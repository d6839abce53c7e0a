#!/bin/bash
# GPU call T (2 GPUs): list path (c5) under torchrun with the host timeline; NCCL dispatch tests
mkdir -p gpurun_out
VF_C5_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload c5 --no-cpu > gpurun_out/r2t_c5_2gpu.json 2> gpurun_out/r2t_c5_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2t_c5_2gpu.json').read().strip().splitlines()[-1]); print('c5 N=2', d['value'], d['ms_per_step'], d['videos_per_sec'])"
grep "c5 trace" gpurun_out/r2t_c5_2gpu.err
timeout 600 python -m pytest tests/test_dispatch_gpu.py -x -q -m gpu 2>&1 | tail -4
