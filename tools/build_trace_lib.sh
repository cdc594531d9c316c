#!/bin/bash
# Development build of the library with the GEMM wait-cycle counters compiled in (-DRPG_GEMM_TRACE); used by
# tools/gemm_trace.py only.  Output: tools/_trace/librpg_b200_trace.so (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_trace
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -DRPG_GEMM_TRACE"
for f in rpg_gemm rpg_gemm_tn rpg_aux rpg_layer rpg_util rpg_attention; do
  /usr/local/cuda/bin/nvcc $FLAGS -c relpose_gnn_b200/csrc/$f.cu -o tools/_trace/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o tools/_trace/librpg_b200_trace.so tools/_trace/*.o -lcudart
rm -f tools/_trace/*.o
echo built tools/_trace/librpg_b200_trace.so
