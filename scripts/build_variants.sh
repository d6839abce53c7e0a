#!/bin/bash
# A/B libraries for the GEMM epilogue experiments: one libvfeat_<tag>.so per variant of csrc/gemm.cu (everything else is
# shared object code).  Select one at run time with VF_LIBVFEAT=<path>.  Usage: scripts/build_variants.sh
set -e
cd "$(dirname "$0")/.."
SRC=video_features_b200/csrc
OBJ=build/obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
pids=()
for f in attn_gemm kernels host clip i3d i3d_kernels raft raft_kernels; do
  if [ ! -f $OBJ/$f.o ] || [ $SRC/$f.cu -nt $OBJ/$f.o ] || [ $SRC/common.cuh -nt $OBJ/$f.o ] || [ $SRC/internal.h -nt $OBJ/$f.o ]; then
    nvcc $FLAGS -c $SRC/$f.cu -o $OBJ/$f.o & pids+=($!)
  fi
done
declare -A V=( [m5g4]="-DVF_EPI_MODE=5 -DVF_EPI_GROUPS=4" [m6g2]="-DVF_EPI_MODE=6 -DVF_EPI_GROUPS=2" \
               [m6g4]="-DVF_EPI_MODE=6 -DVF_EPI_GROUPS=4" [m6g4s6]="-DVF_EPI_MODE=6 -DVF_EPI_GROUPS=4 -DVF_STAGES_256=6" \
               [m5g4s4]="-DVF_EPI_MODE=5 -DVF_EPI_GROUPS=4 -DVF_STAGES_256=4" [m6g3]="-DVF_EPI_MODE=6 -DVF_EPI_GROUPS=3" )
for tag in "${!V[@]}"; do
  nvcc $FLAGS ${V[$tag]} -c $SRC/gemm.cu -o $OBJ/gemm_$tag.o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
for tag in "${!V[@]}"; do
  nvcc -shared -o video_features_b200/libvfeat_$tag.so $OBJ/gemm_$tag.o $OBJ/attn_gemm.o $OBJ/kernels.o $OBJ/host.o $OBJ/clip.o $OBJ/i3d.o \
       $OBJ/i3d_kernels.o $OBJ/raft.o $OBJ/raft_kernels.o
done
ls -la video_features_b200/libvfeat_*.so
