#!/bin/bash
# A/B: checkpoint interval (1 = every state stored, no recompute pass) x gradient zero-fill mode
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
for w in c3 shard; do
for k in 16 1; do
for z in 0 1; do
HBV_B200_FUSED_ZERO=$z python bench.py --workload $w --ckpt $k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${w}_k${k}_z${z}.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_${w}_k${k}_z${z}.json'))
print('$w K=$k zero=$z', 'ms %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
PY
done; done; done
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/bench_c2_b.json; python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_b.json'))
print('c2 ms', d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], 'fwd', d['fwd']['ms_per_step'], d['clocks'], 'e2e', d['e2e']['ms_per_step'])
PY
