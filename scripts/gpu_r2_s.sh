#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/s_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/s_tests.log | head -30
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err ) 2> gpurun_out/s_bench.time
tail -3 gpurun_out/s_bench.err; cat gpurun_out/s_bench.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/s_bench.json'))
print('c2 ms/step', d['ms_per_step'], 'value %.3e'%d['value'], d['run_info'])
print('kernels', {k: round(v,4) for k,v in d['kernel_ms'].items()})
print('e2e', {k: d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['c1_fwd']['value'])
for n, a in d['at_scale'].items():
    if 'error' in a: print(n, a); continue
    print(n, 'ms %.3f fwd-only %.3f' % (a['ms_per_step'], a['fwd_ms_per_step']), {k: round(v, 3) for k, v in a['kernel_ms'].items()},
          'bwd frac %.3f (survey %s) fwd frac %.3f (survey %s)' % (a['roofline']['frac'], a['roofline'].get('frac_survey'), a['roofline_fwd']['frac'], a['roofline_fwd'].get('frac_survey')))
PY
