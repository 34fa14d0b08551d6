#!/bin/bash
# round 2, call G: ncu of the pipelined forward (with state stores) at C4 + GPU tests
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_lean_gpu.py tests/test_dense_gpu.py -m gpu -q > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -4 gpurun_out/g_tests.log
HBV_B200_PIPE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_fwd_.*_kernel -s 2 -c 1 -f -o /tmp/ncu/prof_c4_fwd python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ncu_c4_fwd.log 2>&1
tail -1 gpurun_out/ncu_c4_fwd.log
ncu -i /tmp/ncu/prof_c4_fwd.ncu-rep --page raw --csv > gpurun_out/c4_fwdpipe_raw.csv 2>/dev/null
ncu -i /tmp/ncu/prof_c4_fwd.ncu-rep --page source --csv --print-source sass > gpurun_out/c4_fwdpipe_src.csv 2>/dev/null
ls -la gpurun_out | tail -4
