#!/bin/bash
# ring staging as base + t * stride (one IMAD.WIDE per copy): suite, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/al_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/al_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/al_tests.log | head
timeout 900 python bench.py > gpurun_out/al_bench.json 2> gpurun_out/al_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/al_bench.err
