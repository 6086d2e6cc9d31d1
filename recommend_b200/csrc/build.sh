#!/bin/bash
# Builds libonetrans_sm100.so in-tree (recommend_b200/lib/). nvcc cross-compiles sm_100a without a GPU.
set -e
cd "$(dirname "$0")"
mkdir -p ../lib obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC ${OT_NVCC_EXTRA}"
# objects are rebuilt when a source is newer - and all of them when the flags differ from the last build's (a debug build with
# OT_NVCC_EXTRA must not leave its objects behind for the next plain build)
if [ ! -f obj/.flags ] || [ "$(cat obj/.flags)" != "$FLAGS" ]; then rm -f obj/*.o; echo "$FLAGS" > obj/.flags; fi
pids=()
for f in ot_api ot_gemm ot_ffn_fused ot_wgrad ot_attn_fwd ot_attn_fwd_ws ot_attn_fwd_v2 ot_attn_fwd_v3 ot_attn_fwd_v4 ot_attn_fwd_v5 ot_attn_bwd ot_attn_bwd_fused ot_attn_bwd_v2 ot_attn_cached ot_elementwise ot_optimizer ot_embedding ot_heads ot_metrics; do
  if [ -f $f.cu ]; then
    if [ ! -f obj/$f.o ] || [ $f.cu -nt obj/$f.o ] || [ ot_common.cuh -nt obj/$f.o ] || [ ot_attn.cuh -nt obj/$f.o ] || [ ot_attn_fwd_common.cuh -nt obj/$f.o ] || [ ot_host.h -nt obj/$f.o ] || [ ot_gemm_common.cuh -nt obj/$f.o ] || [ ../../include/onetrans_b200.h -nt obj/$f.o ]; then
      nvcc $FLAGS -c $f.cu -o obj/$f.o &
      pids+=($!)
    fi
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o ../lib/libonetrans_sm100.so obj/*.o
echo "built $(realpath ../lib/libonetrans_sm100.so)"
