#!/bin/bash
# where does the one-warp-CTA + ring form stop winning?  hbv D2 fwd+bwd at several basin counts
for B in 2500 5000 10000 22500; do
for thr in 0 100000000; do
HBV_B200_LEAN_SMALL=$thr timeout 300 python bench.py --workload shard --basins $B --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('B=$B small_thr=$thr', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, 'fwd-only', {k:round(v,3) for k,v in d['fwd']['kernel_ms'].items()})
"
done; done
for thr in 0 100000000; do
HBV_B200_LEAN_SMALL=$thr timeout 300 python scripts/bench_configs.py c4 --steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('c4 small_thr=$thr', round(d['ms_per_step'],2), round(d['fwd_ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})
    except Exception as e: print(l[:200])
"
done
