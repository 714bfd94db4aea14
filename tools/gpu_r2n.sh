#!/bin/bash
# source-level ncu capture of the pipelined TMA column pass (main launch of the key switch)
export HEON_COL_TMA=1 HEON_COL_TMA_TILES=8
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ntt_col_pass_tma_pipe<heon::MapDigitSkip" -s 1 -c 1 -o gpurun_out/r2_colpipe -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_colpipe.log 2>&1
tail -2 gpurun_out/ncu_colpipe.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
