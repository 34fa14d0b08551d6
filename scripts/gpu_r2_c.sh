#!/bin/bash
# round 2, call C: ncu of the C4 recurrence kernels, pipelined vs lean
mkdir -p gpurun_out
for pipe in 1 0; do
HBV_B200_PIPE=$pipe timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 4 -c 2 -f -o gpurun_out/prof_c4_pipe$pipe python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ncu_c4_pipe$pipe.log 2>&1
tail -2 gpurun_out/ncu_c4_pipe$pipe.log
done
