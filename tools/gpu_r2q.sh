#!/bin/bash
# TFHE: parity + workload timing
timeout 600 python -m pytest tests/test_gpu_tfhe.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload M5_tfhe_nand --steps 5 --warmup 3 > gpurun_out/bench_r2c_M5_tfhe_nand.json 2> gpurun_out/bench_r2c_M5.err
tail -2 gpurun_out/bench_r2c_M5.err
python - <<PY
import json
for f in ('bench_r2c_M5_tfhe_nand',):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')][-1])
        print(f, 'value', d.get('value'), 'ms/step', d.get('ms_per_step'), 'ok', d.get('decrypts_correctly'), 'e2e', d.get('e2e',{}).get('value'), 'roof', (d.get('roofline') or {}).get('frac'), [(k['kernel'], round(k['ms_per_step'],2)) for k in d.get('kernels',[])])
    except Exception as e: print(f, 'failed', e)
PY
