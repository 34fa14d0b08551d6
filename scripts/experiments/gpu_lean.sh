#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lean_gpu.py tests/test_dense_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -12
for l in 0 1; do
HBV_B200_LEAN=$l timeout 300 python bench.py --workload shard --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/lean_$l.json 2>gpurun_out/lean.err || tail -5 gpurun_out/lean.err
python - <<PY
import json
d=json.load(open('gpurun_out/lean_$l.json'))
print('shard lean=$l', 'ms %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
PY
HBV_B200_LEAN=$l timeout 200 python scripts/experiments/ab_hbv2.py 22500 730 2>&1 | tail -1
HBV_B200_LEAN=$l timeout 300 python scripts/bench_configs.py c4 --steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('c4 lean=$l', round(d['ms_per_step'],2), round(d['fwd_ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()}, d['checks'])
    except Exception as e: print(l[:200])
"
done
