python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for r in 0 1; do
HBV_B200_RING=$r python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ring_$r.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/ring_$r.json'))
print('RING=$r c2: step %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v,3) for k,v in d['kernel_ms'].items()})
for n,a in d['at_scale'].items(): print('   ', n, 'step %.3f fwd-only %.3f' % (a['ms_per_step'], a['fwd_ms_per_step']), {k: round(v,3) for k,v in a['kernel_ms'].items()})
PY
done
