#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ntt or reference_kernels or alternate" 2>&1 | tail -3
run() { # workload env...
  wl=$1; shift
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m.json 2> gpurun_out/r2m.err
  python - "$wl $*" <<PY
import json,sys
d=json.loads([l for l in open('gpurun_out/r2m.json') if l.startswith('{')][-1])
ks={k['kernel']: round(k['ms_per_op']*1000,1) for k in d['kernels']}
print(sys.argv[1], 'value', round(d['value'],1), 'ntt frac', round(d['roofline_ntt']['frac'],4), 'us/poly', round(d['roofline_ntt']['us_per_limb_poly'],4), 'col', ks.get('ntt_fwd_col_pass'), 'row', ks.get('ntt_fwd_row_pass'))
PY
}
run C3_II HEON_ROW_WALK=0
run C3_II HEON_ROW_WALK=8
run C3_II HEON_ROW_WALK=4
run C3_II HEON_ROW_WALK=2
run C3_I HEON_ROW_WALK=0
run C3_I HEON_ROW_WALK=8
