#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_long_hourly_gpu.py tests/test_parity_gpu.py tests/test_lean_gpu.py tests/test_dense_gpu.py -m gpu -q > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/h_tests.log
grep -E "FAILED|passed|failed|AssertionError:" gpurun_out/h_tests.log | head -40
