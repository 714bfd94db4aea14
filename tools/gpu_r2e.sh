#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ntt_col_pass|k_row_mac|k_modup2_prep" -s 10 -c 8 -o /tmp/prof_r2e -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
ncu -i /tmp/prof_r2e.ncu-rep --page raw --csv > gpurun_out/r2e_raw.csv 2>/dev/null
python tools/summarize_ncu.py full /tmp/prof_r2e.ncu-rep gpurun_out/r2e_summary.csv
cat gpurun_out/r2e_summary.csv
ls -la /tmp/prof_r2e.ncu-rep
