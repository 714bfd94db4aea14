#!/bin/bash
# One GPU-box session: micro-benchmarks, parity tests, smoke, bench (ours + reference).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./tools/microbench > gpurun_out/microbench.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
for w in C3_II C3_I; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  timeout 600 python bench.py --workload $w --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$w.json 2> gpurun_out/bench_ref_$w.err
done
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/microbench.txt
for f in gpurun_out/bench_*.json; do echo $f; head -c 600 $f; echo; done
