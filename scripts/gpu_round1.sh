#!/bin/bash
# GPU box: tests, smoke, bench, ncu launch list, ncu full capture of the recurrence kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 6 -c 3 -o gpurun_out/prof_shard python bench.py --workload shard --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shard.log 2>&1
tail -3 gpurun_out/ncu_shard.log
ls -la gpurun_out
