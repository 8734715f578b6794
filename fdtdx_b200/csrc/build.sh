#!/bin/bash
# Builds the C-ABI shared library in-tree (fdtdx_b200/libfdtdx_b200.so) for sm_100a.
# -fmad=false / IEEE div: every float32 expression rounds like the reference's op-by-op jnp code.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${FDTDX_OUT:-$HERE/../libfdtdx_b200.so}"
OBJ="${FDTDX_OBJ:-${TMPDIR:-/tmp}/fdtdx_b200_obj}"  # objects stay out of the tree (the tree is what travels to the GPU box)
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC ${FDTDX_NVCC_EXTRA:-}"
pids=()
for u in abi yee_E4 yee_E1 yee_H4 yee_H1 yee_E4t yee_H4t yee_E4t64 yee_H4t64 yee_E4tf yee_H4tf; do
  "$NVCC" $FLAGS -c "$HERE/$u.cu" -o "$OBJ/$u.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ/abi.o" "$OBJ/yee_E4.o" "$OBJ/yee_E1.o" "$OBJ/yee_H4.o" "$OBJ/yee_H1.o" "$OBJ/yee_E4t.o" "$OBJ/yee_H4t.o" "$OBJ/yee_E4t64.o" "$OBJ/yee_H4t64.o" "$OBJ/yee_E4tf.o" "$OBJ/yee_H4tf.o"
echo "built $OUT"
