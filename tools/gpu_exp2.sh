#!/bin/bash
mkdir -p gpurun_out
{
echo "== parity (new modup2/moddown kernels)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default lib (minblocks 3)"
python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
echo "== minblocks 4"
HEON_B200_LIB=$PWD/heongpu_b200/lib/libheon_mb4.so python tools/time_ntt.py n16_II_small 37 2>&1 | tail -3
echo "== bench C3_II"
python bench.py --workload C3_II --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k in d['kernels']: print('   %-18s ms/op %.4f share %.3f'%(k['kernel'],k['ms_per_op'],k['share']))"
} > gpurun_out/exp2.txt 2>&1
cat gpurun_out/exp2.txt
