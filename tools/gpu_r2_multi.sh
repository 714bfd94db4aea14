#!/bin/bash
# multi-GPU: pinned-copy ceiling with N GPUs driven concurrently, then the default bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_ceiling.py > gpurun_out/r2_pcie_${N}gpu.json 2> gpurun_out/r2_pcie_${N}gpu.err; cat gpurun_out/r2_pcie_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_${N}gpu_C3_II.json 2> gpurun_out/bench_r2_${N}gpu.err; tail -c 400 gpurun_out/bench_r2_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_${N}gpu_C3_II.json').read().strip().splitlines()[-1]); print('n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'])"
nvidia-smi topo -m | head -12
