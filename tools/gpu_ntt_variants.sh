#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do
  for set in n16_II_small n16_I_small; do
    HEON_NTT_VARIANT=$v python tools/time_ntt.py $set 37 2>&1 | tail -3
  done
  HEON_NTT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "ntt or relin" 2>&1 | tail -2
done > gpurun_out/ntt_variants.txt 2>&1
cat gpurun_out/ntt_variants.txt
