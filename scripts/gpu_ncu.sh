#!/bin/bash
# usage: gpu_ncu.sh <workload> <skip> <count> [kernel regex]  -> gpurun_out/prof_<workload>.ncu-rep
# ncu --set full capture of the recurrence kernels of one bench workload
WL=${1:-shard}; SKIP=${2:-6}; CNT=${3:-3}; RX=${4:-hbv_.*_kernel}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -f -o gpurun_out/prof_$WL python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/ncu_$WL.log 2>&1
tail -2 gpurun_out/ncu_$WL.log
