#!/bin/bash
# Round artefacts: parity, bench (ours + reference), ncu launch list and full captures.
R=${1:-r1}
mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu_$R.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$R.txt
fi
for w in C3_II C3_I; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${R}_$w.json 2> gpurun_out/bench_${R}_$w.err
  timeout 400 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_${R}_ref_$w.json 2> gpurun_out/bench_${R}_ref_$w.err
done
for w in n14_C2 M4_bfv_rot; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_$w.json 2> gpurun_out/bench_${R}_$w.err
  timeout 300 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_${R}_ref_$w.json 2> gpurun_out/bench_${R}_ref_$w.err
done
timeout 120 python tools/time_hoisted.py C3_II 8 4 > gpurun_out/hoisted_$R.txt 2>&1
timeout 120 python tools/time_hoisted.py C3_I 8 2 >> gpurun_out/hoisted_$R.txt 2>&1
# launch list of the default bench command (per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}_C3_II.csv python bench.py --workload C3_II --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# one full capture of every kernel class of the step
# (the .ncu-rep files stay on the box: gpurun_out is capped at 64 MiB; the raw pages come back as CSV)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_row_pass_tma|ntt_col_pass|k_keyswitch_mac|k_modup2|k_moddown2|k_cross_multiply" -s 12 -c 12 -o /tmp/prof_${R}_C3_II -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_${R}_C3_II.ncu-rep --page raw --csv > gpurun_out/prof_${R}_C3_II_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ntt_row_pass_tma|ntt_col_pass" -s 4 -c 2 -o /tmp/prof_${R}_C3_I -f python bench.py --workload C3_I --steps 2 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/ncu_full_I.log 2>&1
ncu -i /tmp/prof_${R}_C3_I.ncu-rep --page raw --csv > gpurun_out/prof_${R}_C3_I_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
tail -2 gpurun_out/pytest_gpu_$R.txt
for f in gpurun_out/bench_${R}_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('e2e',{}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_op') or {}).get('frac'))"; done
cat gpurun_out/hoisted_$R.txt
