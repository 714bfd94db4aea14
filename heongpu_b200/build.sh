#!/bin/bash
# Build libheon_b200.so (sm_100a) in-tree.  No GPU needed (nvcc cross-compiles).
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${HEON_OUT:-libheon_b200.so}
OBJ=${HEON_OBJDIR:-lib}
mkdir -p $OBJ
FLAGS="$HEON_EXTRA -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static"
pids=()
for f in ntt ckks_ops bfv_ops context capi; do
  $NVCC $FLAGS -c csrc/$f.cu -o $OBJ/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC $FLAGS -shared $OBJ/ntt.o $OBJ/ckks_ops.o $OBJ/bfv_ops.o $OBJ/context.o $OBJ/capi.o -o lib/$OUT
echo "built $(pwd)/lib/$OUT"
