#!/bin/bash
# quick GPU check: parity tests + margins (SFU and precise builds), smoke, bench, optional ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== parity margins, default (SFU math) build"; python scripts/parity_report.py 2>&1 | tail -7
if [ -f hydrodl2_b200/lib/libhbv_b200_precise.so ]; then
echo "== parity margins, precise-math build"; HBV_B200_LIB=$PWD/hydrodl2_b200/lib/libhbv_b200_precise.so python scripts/parity_report.py 2>&1 | tail -7
fi
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('c2 ms/step', d['ms_per_step'], 'value', d['value'], 'fwd', d['fwd'])
print('c2 kernels', d['kernel_ms'])
print('e2e', d['e2e']['ms_per_step'])
for n, a in d['at_scale'].items():
    print(n, 'ms %.3f fwd-only %.3f' % (a['ms_per_step'], a['fwd_ms_per_step']), {k: round(v, 3) for k, v in a['kernel_ms'].items()}, 'frac bwd %.3f fwd %.3f' % (a['roofline']['frac'], a['roofline_fwd']['frac']))
PY
if [ "$1" == "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 6 -c 3 -o gpurun_out/prof_shard2 python bench.py --workload shard --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shard2.log 2>&1
tail -2 gpurun_out/ncu_shard2.log
fi
