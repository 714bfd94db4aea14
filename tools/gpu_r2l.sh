#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload M1_bfv_latency --steps 5 --warmup 3 > gpurun_out/bench_r2b_M1.json 2> gpurun_out/bench_r2b_M1.err; tail -c 400 gpurun_out/bench_r2b_M1.err; tail -c 600 gpurun_out/bench_r2b_M1.json
timeout 300 python bench.py --workload C3_II --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3_II value',d['value'],'e2e',d['e2e'])"
# compute-sanitizer: memcheck + racecheck over the NTT tests and the alternate operator paths (TMA transposes,
# mbarrier pipelines, the fused row-pass + inner-product kernel)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "ntt_matches_oracle or alternate_operator_paths or ntt_edge" --timeout 800 > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_sanitizer_memcheck.txt
tail -12 gpurun_out/r2_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "ntt_matches_oracle and (n12_I or n13_II) or (alternate_operator_paths and n13_II)" --timeout 800 > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_sanitizer_racecheck.txt
tail -12 gpurun_out/r2_sanitizer_racecheck.txt
