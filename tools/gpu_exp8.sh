#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -o gpurun_out/prof_ntt_v3 -f python tools/run_ntt.py n16_I_small 3 37 2>&1 | tail -2
