#!/bin/bash
{
echo "== null arithmetic (memory/transposition floor)"
HEON_B200_LIB=$PWD/heongpu_b200/lib/libheon_null.so python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
HEON_B200_LIB=$PWD/heongpu_b200/lib/libheon_null.so HEON_NTT_TMA=0 python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
echo "== real"
python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
python tools/time_ntt.py n16_I_small 148 2>&1 | tail -3
} 2>&1 | grep -E "==|us/poly"
