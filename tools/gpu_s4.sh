#!/bin/bash
tag=${1:-s4}
mkdir -p gpurun_out
{
for w in n14_C2 M4_bfv_rot; do
echo "== $w ours"
timeout 300 python bench.py --workload $w --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${tag}_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d['clocks'], 'launches', d['gpu_launches'])
for k in d['kernels']: print('   %-18s ms/op %.5f share %.3f'%(k['kernel'],k['ms_per_op'],k['share']))
print(d.get('roofline'))"
echo "== $w reference"
timeout 300 python bench.py --workload $w --impl reference --steps 3 2>&1 | tail -1 | tee gpurun_out/${tag}_ref_$w.json | cut -c1-400
done
} > gpurun_out/$tag.txt 2>&1
cat gpurun_out/$tag.txt
