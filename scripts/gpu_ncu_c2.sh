#!/bin/bash
# ncu --set full capture of the recurrence kernels at the C2 (531-basin, latency-bound) size
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 9 -c 3 -f -o gpurun_out/prof_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/ncu_c2_full.log 2>&1
tail -2 gpurun_out/ncu_c2_full.log
ls -la gpurun_out
