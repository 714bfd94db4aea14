#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_bfv_decrypt_level.py -x -q 2>&1 | tail -15
