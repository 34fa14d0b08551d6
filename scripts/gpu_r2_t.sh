#!/bin/bash
# round 2 evidence: launch list of the bench command, ncu --set full of the C2 kernels and of the shard's K = 4 kernels
mkdir -p gpurun_out /tmp/ncu
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-at-scale --no-cpu-baseline > gpurun_out/t_launch.log 2>&1
tail -1 gpurun_out/t_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_.*pipe_kernel -s 9 -c 3 -f -o /tmp/ncu/r02_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/t_ncu_c2.log 2>&1
tail -1 gpurun_out/t_ncu_c2.log
ncu -i /tmp/ncu/r02_c2.ncu-rep --page raw --csv > gpurun_out/r02_c2_raw.csv 2>/dev/null
ncu -i /tmp/ncu/r02_c2.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_c2_src.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:hbv_.*lean_kernel -s 6 -c 3 -f -o /tmp/ncu/r02_shard python bench.py --workload shard --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/t_ncu_shard.log 2>&1
tail -1 gpurun_out/t_ncu_shard.log
ncu -i /tmp/ncu/r02_shard.ncu-rep --page raw --csv > gpurun_out/r02_shard_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
