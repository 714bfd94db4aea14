#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_reference_cpp_tests.py -m gpu -q --timeout 900 -k "benchmarks_and_examples" > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.txt
grep -n "what():\|passed\|failed" gpurun_out/r2j_pytest.txt | head -20
tail -32 gpurun_out/refcpp_benchmark_bfv.txt
tail -20 gpurun_out/refcpp_9_multi_stream_usage_way1.txt
