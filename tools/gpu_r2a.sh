#!/bin/bash
# round-2 first GPU session: parity (all gpu tests), microbenchmarks, PCIe ceiling, baseline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.txt
tail -15 gpurun_out/r2a_pytest.txt
./tools/microbench4 > gpurun_out/r2a_microbench4.txt 2>&1; cat gpurun_out/r2a_microbench4.txt
timeout 120 python tools/pcie_ceiling.py > gpurun_out/r2a_pcie_1gpu.json 2> gpurun_out/r2a_pcie.err; cat gpurun_out/r2a_pcie_1gpu.json
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2a_topo.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_C3_II.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench_C3_II.json
