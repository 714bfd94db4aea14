#!/bin/bash
mkdir -p gpurun_out
{
python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -o gpurun_out/prof_ntt_v2 -f python tools/run_ntt.py n16_I_small 3 37 2>&1 | tail -2
} > gpurun_out/exp5.txt 2>&1
cat gpurun_out/exp5.txt
