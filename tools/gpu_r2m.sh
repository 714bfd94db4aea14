#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.txt
tail -15 gpurun_out/r2m_pytest.txt
