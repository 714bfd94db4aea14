#!/bin/bash
mkdir -p gpurun_out
{
for by in 1 2 4 8; do
echo "== HEON_MAC_BY=$by"
for w in C3_II C3_I; do
HEON_MAC_BY=$by timeout 200 python bench.py --workload $w --steps 6 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w value',round(d['value'],1), [round(k['ms_per_op'],4) for k in d['kernels'] if k['kernel']=='keyswitch_mac'])"
done; done
} > gpurun_out/macby.txt 2>&1
cat gpurun_out/macby.txt
