#!/bin/bash
mkdir -p gpurun_out
for k in 1 2 4; do
HBV_B200_CKPT=$k timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/r_c4_k$k.json 2> gpurun_out/r_c4_k$k.err
python - <<PY
import json
for ln in open('gpurun_out/r_c4_k$k.json'):
    c=json.loads(ln); print('c4 K=$k',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['prefix_bit_exact'])
PY
for B in 531 5000 10000; do
timeout 600 python bench.py --workload shard --basins $B --ckpt $k --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/r_b${B}_k$k.json 2> gpurun_out/r_b${B}_k$k.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/r_b${B}_k$k.json'))
    print('B=$B K=$k ms',round(b['ms_per_step'],3),{kk: round(v,3) for kk,v in b['kernel_ms'].items()})
except Exception as e: print('B=$B K=$k',e)
PY
done
done
