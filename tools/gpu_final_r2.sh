#!/bin/bash
# Last validation session of round 2 (one B200): whole -m gpu suite, default bench in both arms, C3-I, launch list,
# full ncu capture of the kernels added last (k_row_final), memcheck + racecheck over them.  Usage: ./tools/gpu_final_r2.sh [tag]
R=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/${R}_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench_${R}_C3_II.json 2> gpurun_out/bench_${R}_C3_II.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_${R}_ref_C3_II.json 2> gpurun_out/bench_${R}_ref_C3_II.err
timeout 200 python bench.py --workload C3_I --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_C3_I.json 2> gpurun_out/bench_${R}_C3_I.err
timeout 200 python bench.py --workload n14_C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_n14_C2.json 2> gpurun_out/bench_${R}_n14_C2.err
python - <<PY
import json
for f in ['bench_${R}_C3_II', 'bench_${R}_ref_C3_II', 'bench_${R}_C3_I', 'bench_${R}_n14_C2']:
    try:
        d = json.loads([l for l in open('gpurun_out/' + f + '.json') if l.startswith('{')][-1])
        print(f, d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'roofline', (d.get('roofline') or {}).get('frac'),
              'op', (d.get('roofline_op') or {}).get('frac'), 'ntt', (d.get('roofline_ntt') or {}).get('frac'), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}_C3_II.csv python bench.py --workload C3_II --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_${R}_C3_II.csv gpurun_out/${R}_launches_C3_II.csv; cat gpurun_out/${R}_launches_C3_II.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_row_final|k_row_mac|k_moddown2_corr|ntt_col_pass_tma_pipe|k_modup2_fast" -s 10 -c 10 -o /tmp/prof_${R} -f python bench.py --workload C3_II --steps 2 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/summarize_ncu.py full /tmp/prof_${R}.ncu-rep gpurun_out/${R}_ncu_full_C3_II.csv; cat gpurun_out/${R}_ncu_full_C3_II.csv
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "alternate_operator_paths and (env23 or env25 or env27 or env30 or env31)" --timeout 450 > gpurun_out/${R}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/${R}_sanitizer_memcheck.txt
tail -4 gpurun_out/${R}_sanitizer_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "alternate_operator_paths and (env23 or env26 or env28)" --timeout 450 > gpurun_out/${R}_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/${R}_sanitizer_racecheck.txt
tail -4 gpurun_out/${R}_sanitizer_racecheck.txt
