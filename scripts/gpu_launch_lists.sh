#!/bin/bash
# GPU call Q: launch lists (ncu, cold-cache/serialised) of one I3D rgb forward and one RAFT call, with tensor-pipe activity
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
VF_ONCE=1 VF_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2q_i3d_launches.csv python scripts/ncu_i3d_once.py rgb 8 > gpurun_out/r2q_i3d.log 2>&1
tail -2 gpurun_out/r2q_i3d.log
VF_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2q_raft_launches.csv python scripts/ncu_raft_once.py 9 > gpurun_out/r2q_raft.log 2>&1
tail -2 gpurun_out/r2q_raft.log
python scripts/ncu_i3d_once.py rgb 8 | tail -1
python scripts/ncu_i3d_once.py rgb 32 | tail -1
