#!/bin/bash
# Builds the standalone C-ABI probes against the in-tree library (rebuild after any header change).
cd "$(dirname "$0")/../.."
mkdir -p tests/native/bin
for p in probe_gemm probe_attn; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/native/bin/$p tests/native/$p.cu \
    -Lrecommend_b200/lib -lonetrans_sm100 -Xlinker -rpath -Xlinker '$ORIGIN/../../../recommend_b200/lib' || exit 1
done
echo "probes built"
