#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
for f in gpurun_out/bench_r1_n$N.json gpurun_out/bench_r1_n1.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['n_gpus'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['traffic'])"; done
tail -3 gpurun_out/bench_r1_n$N.err
