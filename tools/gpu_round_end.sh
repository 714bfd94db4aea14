#!/bin/bash
# Round-end GPU session (one B200): the whole -m gpu suite, every bench workload in both arms, compute-sanitizer
# over the new kernels.  Outputs under gpurun_out/ (copied to profiles/ by hand).  Usage: ./tools/gpu_round_end.sh [tag]
R=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/${R}_pytest_gpu.txt
for wl in C3_II C3_I n14_C2 M4_bfv_rot M1_bfv_latency M5_tfhe_nand; do
  extra=""; [ "$wl" = "C3_II" ] || extra="--no-cpu-baseline"
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 $extra > gpurun_out/bench_${R}_$wl.json 2> gpurun_out/bench_${R}_$wl.err
  timeout 900 python bench.py --impl reference --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${R}_ref_$wl.json 2> gpurun_out/bench_${R}_ref_$wl.err
  python - <<PY
import json
def last(f):
    try: return json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e: return {'value': None, 'err': str(e)}
a, b = last('gpurun_out/bench_${R}_$wl.json'), last('gpurun_out/bench_${R}_ref_$wl.json')
print('$wl', 'ours', a.get('value'), 'e2e', (a.get('e2e') or {}).get('value'), 'roofline', (a.get('roofline') or {}).get('frac'), 'ntt', (a.get('roofline_ntt') or {}).get('frac'), '| reference', b.get('value'), b.get('reference_modes'))
PY
done
timeout 300 python tools/bench_bsgs.py > gpurun_out/${R}_bsgs_C3_II.json 2> gpurun_out/${R}_bsgs.err; cut -c1-400 gpurun_out/${R}_bsgs_C3_II.json
# compute-sanitizer over the kernels added or changed since the last capture
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tfhe.py tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -q -x -k "tfhe or gate or blind or key_switch or bsgs or accumulate or alternate_ntt_paths or alternate_operator_paths" --timeout 1100 > gpurun_out/${R}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/${R}_sanitizer_memcheck.txt
tail -6 gpurun_out/${R}_sanitizer_memcheck.txt
# racecheck: only tests that run THIS engine's kernels (the reference's own TFHE kernels report shared-memory hazards:
# their SmallForwardNTT runs its last six stages without a barrier)
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tfhe.py tests/test_gpu_parity.py -q -x -k "match_cpu_oracle or gate_errors or (alternate_ntt_paths and (TILES or WALK)) or (alternate_operator_paths and n13_II)" --timeout 1100 > gpurun_out/${R}_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/${R}_sanitizer_racecheck.txt
tail -6 gpurun_out/${R}_sanitizer_racecheck.txt
