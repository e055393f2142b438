#!/usr/bin/env bash
# Builds libodmsspe_b200.so (CUDA kernels + C ABI) for sm_100a, in-tree.
# -fmad=false / -ffp-contract=off: FP64 thermodynamics must round like scalar C without FMA contraction.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${MSSPE_OUT:-$HERE/libodmsspe_b200.so}"
SRCS=("$HERE"/csrc/ctx.cu "$HERE"/csrc/kmer_build.cu "$HERE"/csrc/kmer_build_fast.cu "$HERE"/csrc/select.cu "$HERE"/csrc/select_part.cu "$HERE"/csrc/select_dist.cu "$HERE"/csrc/thal_params.cu "$HERE"/csrc/thal.cu "$HERE"/csrc/filters.cu "$HERE"/csrc/graph.cu "$HERE"/csrc/fasta.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -pthread,-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function ${MSSPE_PTXAS_V:+-Xptxas -v} ${MSSPE_DEFS:-} \
  -shared -o "$OUT" "${SRCS[@]}" -lcudart -ldl
echo "built $OUT"
