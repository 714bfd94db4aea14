#!/bin/bash
# ncu --set full capture of the TFHE blind rotation (one full wave of CTAs: 592 samples)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tfhe_blind_rotate|k_tfhe_keyswitch" -s 2 -c 2 -o gpurun_out/r2_tfhe -f python bench.py --workload M5_tfhe_nand --batch 592 --steps 1 --warmup 3 > gpurun_out/ncu_tfhe.log 2>&1
ncu -i gpurun_out/r2_tfhe.ncu-rep --page raw --csv > gpurun_out/r2_tfhe_raw.csv 2>/dev/null
ls -la gpurun_out/r2_tfhe.ncu-rep
