#!/bin/bash
# why does the one-warp per-step ring forward lose 2.8 ms to the 128-thread chunk ring at C4 once it stores states?
mkdir -p gpurun_out /tmp/ncu
HBV_B200_LEAN_SMALL=400000 timeout 900 ncu --set full --import-source on --clock-control none -k regex:hbv_fwd_lean_kernel -s 2 -c 1 -f -o /tmp/ncu/aq_c4 python scripts/bench_configs.py c4 --steps 1 > gpurun_out/aq_ncu.log 2>&1
ncu -i /tmp/ncu/aq_c4.ncu-rep --page raw --csv > gpurun_out/aq_c4_raw.csv 2>/dev/null
ncu -i /tmp/ncu/aq_c4.ncu-rep --page source --csv --print-source sass > gpurun_out/aq_c4_source.csv 2>/dev/null
ls -la gpurun_out/aq_*
