#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 4 -o gpurun_out/prof_ntt_v0 -f python tools/run_ntt.py n16_II_small 3 37 > gpurun_out/ncu_ntt.log 2>&1
tail -5 gpurun_out/ncu_ntt.log
