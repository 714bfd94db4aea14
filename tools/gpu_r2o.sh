#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_client_side.py -m gpu -q -x --timeout 900 > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.txt
tail -5 gpurun_out/r2o_pytest.txt
for cfg in "HEON_COL_TMA=0" "HEON_COL_TMA=1"; do
for w in C3_II C3_I; do
env $cfg timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2o_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg $w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ntt frac',round(d['roofline_ntt']['frac'],3), round(d['roofline_ntt']['us_per_limb_poly'],3))
for k in d['kernels']: print('   ',k['kernel'],round(k['ms_per_op']*1000,1),'us/op')
"
done; done
