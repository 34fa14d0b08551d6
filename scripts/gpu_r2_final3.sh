#!/bin/bash
# closing bench line + launch list of the bench command on the final build
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-at-scale --no-cpu-baseline > gpurun_out/t_launch.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
