#!/bin/bash
mkdir -p gpurun_out
{
echo "== FP64 quotient on"
python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
echo "== FP64 quotient off"
HEON_NTT_FP64=0 python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
./tools/gpu_bench_both.sh
} > gpurun_out/exp6.txt 2>&1
cat gpurun_out/exp6.txt
