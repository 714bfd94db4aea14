#!/bin/bash
# session-2 experiment driver: NTT timing, parity, bench summaries -> gpurun_out/$1.txt
tag=${1:-s2}
mkdir -p gpurun_out
{
echo "== NTT timing"
python tools/time_ntt.py C3_II 8 2>&1 | tail -3
python tools/time_ntt.py C3_I 8 2>&1 | tail -3
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"
for w in C3_II C3_I; do
python bench.py --workload $w --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d['clocks'])
for k in d['kernels']: print('   %-18s ms/op %.4f share %.3f alg_gbs %s'%(k['kernel'],k['ms_per_op'],k['share'],k['alg_gbs']))"
done
} > gpurun_out/$tag.txt 2>&1
cat gpurun_out/$tag.txt
