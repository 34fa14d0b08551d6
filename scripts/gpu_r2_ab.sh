#!/bin/bash
# deep-prefetch K1s confirmed: full suite, default bench, C4 timing, ncu (raw + source pages) of K1s / K2s at C4
mkdir -p gpurun_out /tmp/ncu
( time timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ab_tests.log 2>&1 ) 2> gpurun_out/ab_tests.time
echo "tests rc=$?" >> gpurun_out/ab_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ab_tests.log | head; head -3 gpurun_out/ab_tests.time
timeout 900 python bench.py > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/ab_bench.err
timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ab_c4.json 2> gpurun_out/ab_c4.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hbv_.*_lean_kernel -s 4 -c 2 -f -o /tmp/ncu/ab_c4 python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ab_ncu_c4.log 2>&1
ncu -i /tmp/ncu/ab_c4.ncu-rep --page raw --csv > gpurun_out/ab_c4_raw.csv 2>/dev/null
ncu -i /tmp/ncu/ab_c4.ncu-rep --page source --csv --print-source sass > gpurun_out/ab_c4_source.csv 2>/dev/null
ls -la gpurun_out/ab_* | head -20
