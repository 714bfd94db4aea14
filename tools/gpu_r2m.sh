#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_client_side.py -x -q -m gpu 2>&1 | tail -3
run() { # workload env...
  wl=$1; shift
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m.json 2> gpurun_out/r2m.err
  python - "$wl $*" <<PY
import json,sys
d=json.loads([l for l in open('gpurun_out/r2m.json') if l.startswith('{')][-1])
ks={k['kernel']: round(k.get('ms_per_op',0)*1000,1) for k in d.get('kernels',[]) if 'ms_per_op' in k}
print(sys.argv[1], 'value', round(d['value'],1), d.get('latency_us_per_op'), ks)
PY
}
run C3_II HEON_MODUP_DOUBLES=0
run C3_II HEON_MODUP_DOUBLES=1
run M1_bfv_latency HEON_ROW_WALK=-1
run M1_bfv_latency HEON_ROW_WALK=0
run M4_bfv_rot HEON_ROW_WALK=-1
