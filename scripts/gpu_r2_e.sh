#!/bin/bash
# round 2, call E: ncu of the C4 recurrence kernels (pipe vs lean); exported to CSV on the box (the reports exceed the copy-back limit)
mkdir -p gpurun_out /tmp/ncu
rm -f gpurun_out/*.ncu-rep
for pipe in 1 0; do
HBV_B200_PIPE=$pipe timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 4 -c 2 -f -o /tmp/ncu/prof_c4_pipe$pipe python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ncu_c4_pipe$pipe.log 2>&1
tail -1 gpurun_out/ncu_c4_pipe$pipe.log
ncu -i /tmp/ncu/prof_c4_pipe$pipe.ncu-rep --page raw --csv > gpurun_out/c4_pipe${pipe}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/prof_c4_pipe$pipe.ncu-rep --page source --csv --print-source sass > gpurun_out/c4_pipe${pipe}_src.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
