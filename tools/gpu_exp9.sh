#!/bin/bash
for l in libheon_mb1.so libheon_mb2.so libheon_mb3.so; do
echo "== $l non-persistent"
HEON_B200_LIB=$PWD/heongpu_b200/lib/$l python tools/time_ntt.py n16_I_small 148 2>&1 | grep us/poly
echo "== $l persistent"
HEON_NTT_PERSISTENT=1 HEON_B200_LIB=$PWD/heongpu_b200/lib/$l python tools/time_ntt.py n16_I_small 148 2>&1 | grep us/poly | head -1
done
