#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_reference_cpp_tests.py tests/test_client_side.py -m gpu -q --timeout 900 > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.txt
tail -40 gpurun_out/r2i_pytest.txt
tail -30 gpurun_out/refcpp_benchmark_ckks.txt
