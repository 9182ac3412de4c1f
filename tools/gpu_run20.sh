#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -m gpu -s > gpurun_out/r02_pytest_20.log 2>&1
echo "pytest rc=$?"; grep -n "kink tensor\|passed\|failed\|FAILED\|AssertionError" gpurun_out/r02_pytest_20.log | head -60
