#!/bin/bash
mkdir -p gpurun_out
for fz in 1 0; do
HBV_B200_FUSED_ZERO=$fz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/n_bench_fz$fz.json 2> gpurun_out/n_bench_fz$fz.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/n_bench_fz$fz.json'))
    print('fz=$fz c2 ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],{k: round(v,4) for k,v in b['kernel_ms'].items()})
except Exception as e: print('c2',e)
PY
done
