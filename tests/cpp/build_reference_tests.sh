#!/bin/bash
# Compiles the reference's OWN test sources (/root/reference/test/*.cpp), UNMODIFIED and where they lie,
# against this repository's class layer (heongpu_b200/include/heongpu/heongpu.hpp + libheon_b200.so) and
# a GoogleTest stand-in.  Outputs go to tests/cpp/_bin/ (git-ignored, shipped to the GPU box by gpurun).
# Only runs where /root/reference exists (the build container).
set -e
cd "$(dirname "$0")"
REF=${REF:-/root/reference}
ROOT=$(cd ../.. && pwd)
mkdir -p _bin
[ -d "$REF/test" ] || { echo "no reference tree: skipping"; exit 0; }
TESTS=${TESTS:-"test_ckks_encoding test_ckks_encryption test_ckks_addition test_ckks_multiplication test_ckks_relinearization test_ckks_rotation_method_1 test_ckks_rotation_method_2 test_bfv_encoding test_bfv_encryption test_bfv_addition test_bfv_multiplication test_bfv_relinearization test_bfv_rotation_method_1 test_bfv_rotation_method_2 test_tfhe_gate_boot"}
pids=()
for t in $TESTS; do
  ( g++ -std=c++17 -O1 -w -I shim -I "$ROOT/heongpu_b200/include" -I /usr/local/cuda/include \
      "$REF/test/$t.cpp" -o _bin/$t \
      -L "$ROOT/heongpu_b200/lib" -lheon_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../../../heongpu_b200/lib' \
      && echo "built $t" ) 2> _bin/$t.log || echo "FAILED to build $t (see tests/cpp/_bin/$t.log)" &
  pids+=($!)
done
# the reference's benchmarks and basic examples (multi-stream OpenMP usage included), same rule: unmodified
EXTRA=${EXTRA:-"benchmark/benchmark_ckks benchmark/benchmark_bfv example/basic/1_basic_bfv example/basic/2_basic_ckks example/basic/3_basic_memorypool_config example/basic/4_switchkey_methods_bfv example/basic/5_switchkey_methods_ckks example/basic/8_default_stream_usage example/basic/9_multi_stream_usage_way1 example/basic/10_multi_stream_usage_way2 example/basic/11_basic_bfv_logic example/basic/12_basic_ckks_logic example/basic/13_bfv_serialization example/basic/14_ckks_serialization example/basic/15_basic_tfhe"}
for e in $EXTRA; do
  b=$(basename $e)
  ( g++ -std=c++17 -O1 -w -fopenmp -I shim -I "$ROOT/heongpu_b200/include" -I /usr/local/cuda/include \
      "$REF/$e.cpp" -o _bin/$b \
      -L "$ROOT/heongpu_b200/lib" -lheon_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../../../heongpu_b200/lib' \
      && echo "built $b" ) 2> _bin/$b.log || echo "FAILED to build $b (see tests/cpp/_bin/$b.log)" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls _bin
