#!/bin/bash
tag=${1:-s3}
mkdir -p gpurun_out
{
echo "== parity"
timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -5
echo "== NTT timing (pipe, group sweep)"
for g in 32 48 64; do echo "G=$g"; HEON_NTT_GROUP=$g timeout 120 python tools/time_ntt.py C3_II 8 2>&1 | grep -v INTT | tail -2; done
HEON_NTT_GROUP=48 timeout 120 python tools/time_ntt.py C3_I 8 2>&1 | tail -3
} > gpurun_out/$tag.txt 2>&1
cat gpurun_out/$tag.txt
