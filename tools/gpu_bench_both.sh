#!/bin/bash
mkdir -p gpurun_out
for w in C3_II C3_I; do
python bench.py --workload $w --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1),'ms/op',round(1000/d['value'],3),'e2e',round(d['e2e']['value'],1), 'kernel-sum ms/op', round(sum(k['ms_per_op'] for k in d['kernels']),3))"
done
timeout 300 python -m pytest tests/test_class_layer.py -q 2>&1 | tail -2
