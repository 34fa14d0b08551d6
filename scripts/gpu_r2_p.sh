#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_reference_on_gpu.py -m gpu -q > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/p_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/p_tests.log | head -30
