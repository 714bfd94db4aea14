#!/bin/bash
# compute-sanitizer memcheck over the C++ class-layer tests (host-resident keys, BSGS, serialization, logic gates, TFHE)
L=$PWD/heongpu_b200/lib
for t in host_keys_test bsgs_test serialization_test; do
  g++ -std=c++17 -O1 -I heongpu_b200/include -I/usr/local/cuda/include tests/cpp/$t.cpp -o /tmp/$t -L $L -lheon_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$L || exit 1
done
: > gpurun_out/r2f_sanitizer_cpp.txt
for exe in /tmp/host_keys_test /tmp/bsgs_test /tmp/serialization_test tests/cpp/_bin/11_basic_bfv_logic tests/cpp/_bin/15_basic_tfhe tests/cpp/_bin/14_ckks_serialization; do
  echo "=== $exe" >> gpurun_out/r2f_sanitizer_cpp.txt
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 $exe > /tmp/san.log 2>&1; rc=$?
  grep -E "ERROR SUMMARY|OK|FAILED|Invalid|Error" /tmp/san.log | tail -4 >> gpurun_out/r2f_sanitizer_cpp.txt
  echo "exit $rc" >> gpurun_out/r2f_sanitizer_cpp.txt
done
cat gpurun_out/r2f_sanitizer_cpp.txt
