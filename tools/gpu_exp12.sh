#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
./tools/gpu_bench_both.sh 2>&1 | grep value
python bench.py --workload C3_I --steps 10 --batch 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C3_I batch 8 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k in d['kernels']: print('   %-18s ms/op %.4f share %.3f'%(k['kernel'],k['ms_per_op'],k['share']))"
