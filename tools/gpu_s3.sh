#!/bin/bash
tag=${1:-s3}
mkdir -p gpurun_out
{
echo "== parity"
timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -5
echo "== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ntt16_fwd_pipe" -s 3 -c 1 -o gpurun_out/prof_$tag -f python tools/time_ntt.py C3_II 8 > gpurun_out/ncu_$tag.log 2>&1; tail -2 gpurun_out/ncu_$tag.log
} > gpurun_out/$tag.txt 2>&1
cat gpurun_out/$tag.txt
