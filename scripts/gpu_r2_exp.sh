#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do
HBV_B200_LIB=$PWD/hydrodl2_b200/lib/libhbv_exp$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/exp_ck$v.json 2> gpurun_out/exp_ck$v.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/exp_ck$v.json'))
    print('exp=$v', {k: round(v,4) for k,v in b['kernel_ms'].items()})
except Exception as e: print('exp=$v',e)
PY
done
