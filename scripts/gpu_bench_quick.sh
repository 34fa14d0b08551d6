#!/bin/bash
# tests (optional: arg "notest" skips) + one default bench run, per-kernel summary
mkdir -p gpurun_out
[ "$1" == "notest" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bq.json 2>gpurun_out/bq.err || tail -5 gpurun_out/bq.err
python - <<PY
import json
d=json.load(open('gpurun_out/bq.json'))
print(d['config']['launch']); print('c2: step %.3f eager %.3f fwd-only %.3f e2e %.2f' % (d['ms_per_step'], d['config']['eager_ms_per_step'], d['fwd']['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v,3) for k,v in d['kernel_ms'].items()}, 'clk', d['clocks']['sm_mhz'])
for n,a in d['at_scale'].items(): print('   ', n, 'step %.3f fwd-only %.3f' % (a['ms_per_step'], a['fwd_ms_per_step']), {k: round(v,3) for k,v in a['kernel_ms'].items()}, 'frac bwd %.3f fwd %.3f' % (a['roofline']['frac'], a['roofline_fwd']['frac']))
PY
