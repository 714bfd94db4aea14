#!/bin/bash
# multi-GPU: pinned-copy ceiling with N GPUs driven concurrently, then the default bench and the TFHE workload under torchrun
N=${1:-2}
R=${2:-r2e}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_ceiling.py > gpurun_out/${R}_pcie_${N}gpu.json 2> gpurun_out/${R}_pcie_${N}gpu.err; tail -c 600 gpurun_out/${R}_pcie_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${R}_${N}gpu_C3_II.json 2> gpurun_out/bench_${R}_${N}gpu.err; tail -c 300 gpurun_out/bench_${R}_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload M5_tfhe_nand --steps 5 --warmup 3 > gpurun_out/bench_${R}_${N}gpu_M5_tfhe_nand.json 2> gpurun_out/bench_${R}_${N}gpu_M5.err; tail -c 300 gpurun_out/bench_${R}_${N}gpu_M5.err
python - <<PY
import json
for f in ('bench_${R}_${N}gpu_C3_II', 'bench_${R}_${N}gpu_M5_tfhe_nand'):
    d=json.loads([l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')][-1]); print(f, 'n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'], d.get('clocks'))
PY
