#!/bin/bash
# Builds the C-ABI shared library in-tree (fdtdx_b200/libfdtdx_b200.so) for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libfdtdx_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -shared ${FDTDX_NVCC_EXTRA:-} \
  -o "$OUT" "$HERE/abi.cu"
echo "built $OUT"
