#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.txt
tail -5 gpurun_out/r2d_pytest.txt
for cfg in "HEON_MODUP_FUSED=1" "HEON_MODUP_FUSED=1 HEON_COL_THREADS=256"; do
  for w in C3_II; do
    echo "== $cfg $w"
    env $cfg timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2d_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
for k in d['kernels']: print('   ',k['kernel'],round(k['ms_per_op']*1000,1),'us/op')
"
  done
done
