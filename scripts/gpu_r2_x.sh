#!/bin/bash
mkdir -p gpurun_out
for bz in 1; do
HBV_B200_BULK_ZERO=$bz timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/x_bench_bz$bz.json 2> gpurun_out/x_bench_bz$bz.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/x_bench_bz$bz.json'))
    print('bulk_zero=$bz c2 ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],{k: round(v,4) for k,v in b['kernel_ms'].items()}, 'e2e', b['e2e']['ms_per_step'], b['e2e']['host_gradient_equals_dense_device_gradient'])
except Exception as e: print('c2',e)
PY
tail -2 gpurun_out/x_bench_bz$bz.err
done
