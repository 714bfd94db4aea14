#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for mb in 100000 16 32 48 64 96; do
echo "== window $mb MB"
HEON_KS_WINDOW_MB=$mb ./tools/gpu_bench_both.sh 2>&1 | grep value
done
