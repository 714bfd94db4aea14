#!/bin/bash
# Closing session of round 2 (one B200): whole -m gpu suite, default bench in both arms, every other workload of this arm.
R=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/${R}_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench_${R}_C3_II.json 2> gpurun_out/bench_${R}_C3_II.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_${R}_ref_C3_II.json 2> gpurun_out/bench_${R}_ref_C3_II.err
for wl in C3_I n14_C2 M4_bfv_rot M1_bfv_latency M5_tfhe_nand; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_$wl.json 2> gpurun_out/bench_${R}_$wl.err
done
python - <<PY
import json
for f in ['bench_${R}_C3_II', 'bench_${R}_ref_C3_II'] + ['bench_${R}_' + w for w in 'C3_I n14_C2 M4_bfv_rot M1_bfv_latency M5_tfhe_nand'.split()]:
    try:
        d = json.loads([l for l in open('gpurun_out/' + f + '.json') if l.startswith('{')][-1])
        print(f, d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'roofline', (d.get('roofline') or {}).get('frac'),
              'op', (d.get('roofline_op') or {}).get('frac'), 'ntt', (d.get('roofline_ntt') or {}).get('frac'), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
