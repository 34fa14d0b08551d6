#!/bin/bash
# stored-state layout x forward form at C4: plane / warp-major layout, 128-thread chunk ring / one-warp per-step ring
mkdir -p gpurun_out
for v in base exp2; do
for small in -1 400000; do
lib=$PWD/hydrodl2_b200/lib/libhbv_b200_$v.so; [ $v = base ] && lib=$PWD/hydrodl2_b200/lib/libhbv_b200.so
HBV_B200_LEAN_SMALL=$small HBV_B200_LIB=$lib timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ap_c4.json 2> gpurun_out/ap_c4.err
python - <<PY
import json
for ln in open('gpurun_out/ap_c4.json'):
    c=json.loads(ln); print('c4 $v lean_small=$small',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['finite'])
PY
done
done
