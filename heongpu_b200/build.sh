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
SRCS="ntt ntt_maps ntt_skip ntt_modup1 ntt_modup2 ntt_divround rowmac modup2col ckks_ops bfv_ops client hostpipe tfhe context capi"
for f in $SRCS; do
  $NVCC $FLAGS -c csrc/$f.cu -o $OBJ/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/$f.o"; done
$NVCC $FLAGS -shared $OBJS -o lib/$OUT -lz
echo "built $(pwd)/lib/$OUT"
