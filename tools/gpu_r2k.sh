#!/bin/bash
mkdir -p gpurun_out
for w in C3_II C3_I; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r2a_$w.json 2> gpurun_out/bench_r2a_$w.err; tail -c 300 gpurun_out/bench_r2a_$w.err
  timeout 400 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2a_ref_$w.json 2> gpurun_out/bench_r2a_ref_$w.err; tail -c 300 gpurun_out/bench_r2a_ref_$w.err
done
for w in n14_C2 M4_bfv_rot M1_bfv_latency; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_$w.json 2> gpurun_out/bench_r2a_$w.err; tail -c 300 gpurun_out/bench_r2a_$w.err
  timeout 600 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2a_ref_$w.json 2> gpurun_out/bench_r2a_ref_$w.err; tail -c 300 gpurun_out/bench_r2a_ref_$w.err
done
for f in gpurun_out/bench_r2a_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','latency_us_per_op')}, 'e2e',d.get('e2e',{}).get('value'), d.get('e2e',{}).get('equals_device_path'), 'roof',(d.get('roofline') or {}).get('kernel'),(d.get('roofline') or {}).get('frac'), 'ntt',(d.get('roofline_ntt') or {}).get('frac'), 'op',(d.get('roofline_op') or {}).get('frac'), d.get('reference_modes'))
except Exception as e: print('ERR',e)
"; done
