#!/bin/bash
HEON_ROW_MAC_OVERLAP=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -x -q -m gpu -k "reference_kernels or alternate_operator or bsgs" 2>&1 | tail -3
run() { # workload env...
  wl=$1; shift
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m.json 2> gpurun_out/r2m.err
  python - "$wl $*" <<PY
import json,sys
d=json.loads([l for l in open('gpurun_out/r2m.json') if l.startswith('{')][-1])
ks={k['kernel']: round(k['ms_per_op']*1000,1) for k in d['kernels']}
print(sys.argv[1], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'row_mac(profile sum)', ks.get('keyswitch_row_mac'))
PY
}
run C3_II HEON_ROW_MAC_OVERLAP=0
run C3_II HEON_ROW_MAC_OVERLAP=1
run C3_II HEON_ROW_MAC_OVERLAP=2
run C3_II HEON_ROW_MAC_OVERLAP=3
