#!/bin/bash
# full GPU test suite + default bench line after the pipelined column pass became the default
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.txt
python bench.py > gpurun_out/bench_r2c_C3_II.json 2> gpurun_out/bench_r2c_C3_II.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2c_C3_II.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'ntt', d['roofline_ntt']['frac'], 'clocks', d['clocks'])
for k in d['kernels']: print('   ', k['kernel'], round(k['ms_per_op']*1000,1), k.get('alg_gbs'))
PY
