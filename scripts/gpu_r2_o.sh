#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_reference_on_gpu.py tests/test_hbv_adj_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/o_tests.log
grep -E "FAILED|passed|failed|Error:|Newton schedules" gpurun_out/o_tests.log | head -30
( time timeout 1500 python scripts/cpu_slices.py --basins 256 > gpurun_out/o_cpu_slices.json 2> gpurun_out/o_cpu_slices.err ) 2> gpurun_out/o_cpu_slices.time
cat gpurun_out/o_cpu_slices.json | head -60; tail -3 gpurun_out/o_cpu_slices.err; cat gpurun_out/o_cpu_slices.time
