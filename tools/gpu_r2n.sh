#!/bin/bash
# source-level ncu capture of the TFHE blind rotation (one full wave of CTAs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tfhe_blind_rotate" -s 1 -c 1 -o gpurun_out/r2_tfhe_br -f python bench.py --workload M5_tfhe_nand --batch 592 --steps 1 --warmup 3 > gpurun_out/ncu_tfhe.log 2>&1
tail -2 gpurun_out/ncu_tfhe.log | cut -c1-200
ls -la gpurun_out/r2_tfhe_br.ncu-rep
