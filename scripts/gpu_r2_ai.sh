#!/bin/bash
# K3^T ring form + pair-routing pad rows: full suite, default bench (at_scale: c5), c4 launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ai_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ai_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ai_tests.log | head
timeout 900 python bench.py > gpurun_out/ai_bench.json 2> gpurun_out/ai_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/ai_bench.err
HBV_B200_RING=0 timeout 600 python scripts/bench_configs.py c5 --steps 5 > gpurun_out/ai_c5_ring0.json 2> gpurun_out/ai_c5_ring0.err
timeout 600 python scripts/bench_configs.py c5 --steps 5 > gpurun_out/ai_c5.json 2> gpurun_out/ai_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ai_c4_launches.csv python scripts/bench_configs.py c4 --steps 1 > /dev/null 2>&1
