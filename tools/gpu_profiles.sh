#!/bin/bash
# Round artefacts: parity, bench (ours + reference), ncu launch list and full captures.
R=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$R.txt
for w in C3_II C3_I; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${R}_$w.json 2> gpurun_out/bench_${R}_$w.err
  timeout 600 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_${R}_ref_$w.json 2> gpurun_out/bench_${R}_ref_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${R}_C3_II.csv python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ntt_row_pass_tma|ntt_col_pass|k_keyswitch_mac|k_modup2" -s 8 -c 8 -o gpurun_out/prof_${R}_C3_II -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/pytest_gpu_$R.txt
for f in gpurun_out/bench_${R}_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('e2e',{}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_op') or {}).get('frac'))"; done
