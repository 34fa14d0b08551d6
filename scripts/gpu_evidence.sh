#!/bin/bash
# GPU evidence pass: parity tests, smoke, bench line, ncu launch list of the bench command,
# one ncu --set full capture of the HBV kernels on the c3 (HBM-bound) and shard (issue-bound) workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2.json'))
print('c2 ms/step', d['ms_per_step'], 'value', d['value'], 'fwd', d['fwd'])
print('c2 kernels', d['kernel_ms'])
print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline'])
for n, a in d['at_scale'].items():
    print(n, 'ms %.3f fwd-only %.3f' % (a['ms_per_step'], a['fwd_ms_per_step']), {k: round(v, 3) for k, v in a['kernel_ms'].items()}, 'frac bwd %.3f fwd %.3f' % (a['roofline']['frac'], a['roofline_fwd']['frac']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/ncu_launch.log 2>&1
for w in c2 c3 shard; do
ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 6 -c 3 -f -o gpurun_out/prof_$w python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/ncu_$w.log 2>&1
tail -2 gpurun_out/ncu_$w.log
done
ls -la gpurun_out
