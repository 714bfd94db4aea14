#!/bin/bash
# Build libheon_b200.so (sm_100a) in-tree.  No GPU needed (nvcc cross-compiles).
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static"
pids=()
for f in ntt ckks_ops context capi; do
  $NVCC $FLAGS -c csrc/$f.cu -o lib/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC $FLAGS -shared lib/ntt.o lib/ckks_ops.o lib/context.o lib/capi.o -o lib/libheon_b200.so
echo "built $(pwd)/lib/libheon_b200.so"
