#!/bin/bash
# adjoint instruction trims (in-place static sums, streamflow-only cotangent): suite, bench, c4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ak_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ak_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ak_tests.log | head
timeout 900 python bench.py > gpurun_out/ak_bench.json 2> gpurun_out/ak_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/ak_bench.err
python scripts/parity_report.py > gpurun_out/ak_parity_report.txt 2> gpurun_out/ak_parity.err; tail -2 gpurun_out/ak_parity.err
