#!/bin/bash
# Builds libfsnerf_b200.so (sm_100a only) next to the Python package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libfsnerf_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="api.cu rays.cu composite.cu occgrid.cu mlp_program.cu mlp_fwd2.cu mlp_bwd2.cu"
cd "$HERE"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared ${FSNERF_NVCC_EXTRA} -o "$OUT" $SRCS
echo "built $OUT"
