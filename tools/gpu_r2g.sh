#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_client_side.py tests/test_reference_cpp_tests.py -m gpu -q --timeout 600 > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.txt
tail -80 gpurun_out/r2g_pytest.txt
