#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_client_side.py -m gpu -q --timeout 300 > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.txt
tail -60 gpurun_out/r2f_pytest.txt
