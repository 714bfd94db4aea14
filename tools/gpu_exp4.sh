#!/bin/bash
mkdir -p gpurun_out
{
echo "== TMA row pass"
python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
python tools/time_ntt.py n16_I_small 37 2>&1 | tail -3
echo "== LSU row pass"
HEON_NTT_TMA=0 python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"
for w in C3_II C3_I; do
python bench.py --workload $w --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k in d['kernels']: print('   %-18s ms/op %.4f share %.3f'%(k['kernel'],k['ms_per_op'],k['share']))"
done
} > gpurun_out/exp4.txt 2>&1
cat gpurun_out/exp4.txt
