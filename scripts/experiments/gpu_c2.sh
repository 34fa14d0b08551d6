#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for r in 1 ""; do
if [ -n "$r" ]; then export HBV_B200_RING=$r; else unset HBV_B200_RING; fi
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-at-scale 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('RING=$r', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, 'fwd-only', round(d['fwd']['ms_per_step'],3), d['fwd']['kernel_ms'])
"
done
unset HBV_B200_RING
ncu --set full --clock-control none --import-source on -k regex:hbv_.*lean_kernel -s 4 -c 2 -f -o gpurun_out/prof_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
