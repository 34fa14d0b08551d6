#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for z in 0 ""; do
if [ -n "$z" ]; then export HBV_B200_FUSED_ZERO=$z; else unset HBV_B200_FUSED_ZERO; fi
for w in c2 shard; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w fused_zero=$z', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})
"
done; done
