#!/bin/bash
# closing bench line + ncu of the shard / C5 kernels with a kept gradient plane
mkdir -p gpurun_out /tmp/ncu
timeout 900 python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:hbv_.*lean_kernel -s 9 -c 3 -f -o /tmp/ncu/r02_shard python bench.py --workload shard --steps 2 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/t_ncu_shard.log 2>&1
ncu -i /tmp/ncu/r02_shard.ncu-rep --page raw --csv > gpurun_out/r02_shard_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:hbv_adj_.*_kernel -s 6 -c 2 -f -o /tmp/ncu/r02_c5 python scripts/bench_configs.py c5 --steps 2 > gpurun_out/t_ncu_c5.log 2>&1
ncu -i /tmp/ncu/r02_c5.ncu-rep --page raw --csv > gpurun_out/r02_c5_raw.csv 2>/dev/null
ls -la gpurun_out/r02_*raw.csv gpurun_out/r02_bench_c2.json
