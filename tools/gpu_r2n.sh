#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.txt
tail -5 gpurun_out/r2n_pytest.txt
for w in C3_II C3_I; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2n_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ntt frac',d['roofline_ntt']['frac'], d['roofline_ntt']['us_per_limb_poly'])
for k in d['kernels']: print('   ',k['kernel'],round(k['ms_per_op']*1000,1),'us/op')
"
done
