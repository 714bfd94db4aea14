#!/bin/bash
# Round artefacts: ncu launch list of the default bench command and one full capture per kernel class.
R=${1:-r2}
mkdir -p gpurun_out
# launch list (per-launch times are cold-cache and serialised: compare SHARES with kernels[].share of the bench line)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}_C3_II.csv python bench.py --workload C3_II --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_${R}_C3_II.csv gpurun_out/${R}_launches_C3_II.csv; cat gpurun_out/${R}_launches_C3_II.csv
# one full capture of every kernel class of a step (batch 4 keeps the report small)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_row_mac|ntt_col_pass|ntt_row_pass_tma|k_modup2|k_moddown2|k_cross_multiply|k_stash" -s 16 -c 16 -o /tmp/prof_${R}_C3_II -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/summarize_ncu.py full /tmp/prof_${R}_C3_II.ncu-rep gpurun_out/${R}_ncu_full_C3_II.csv; cat gpurun_out/${R}_ncu_full_C3_II.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_row_mac|ntt_col_pass" -s 4 -c 4 -o /tmp/prof_${R}_C3_I -f python bench.py --workload C3_I --steps 2 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/ncu_full_I.log 2>&1
python tools/summarize_ncu.py full /tmp/prof_${R}_C3_I.ncu-rep gpurun_out/${R}_ncu_full_C3_I.csv; cat gpurun_out/${R}_ncu_full_C3_I.csv
# a small report of the two dominant kernels comes home for the source-level view
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_row_mac" -s 2 -c 1 -o gpurun_out/${R}_rowmac -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
