#!/bin/bash
# chunk-ring K1s (8 basins per CTA) + register-window pair routing + tap-parallel gamma kernels: suite, bench, C4 launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/af_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/af_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/af_tests.log | head
timeout 900 python bench.py > gpurun_out/af_bench.json 2> gpurun_out/af_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/af_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/af_c4_launches.csv python scripts/bench_configs.py c4 --steps 1 > /dev/null 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
