#!/bin/bash
# K1s per-step ring with the pops one step ahead: parity + timing (C4 both forward forms, hbv grids, bench)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ar_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ar_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ar_tests.log | head
for small in -1 400000; do
HBV_B200_LEAN_SMALL=$small timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ar_c4.json 2> gpurun_out/ar_c4.err
python - <<PY
import json
for ln in open('gpurun_out/ar_c4.json'):
    c=json.loads(ln); print('c4 lean_small=$small',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['prefix_bit_exact'])
PY
done
for B in 1000 2000 2500; do
timeout 600 python bench.py --workload shard --basins $B --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/ar_b$B.json 2> gpurun_out/ar_b$B.err
python - <<PY
import json
b=json.load(open('gpurun_out/ar_b$B.json'))
print('hbv B=$B ms',round(b['ms_per_step'],3),{kk: round(v,3) for kk,v in b['kernel_ms'].items()}, 'fwd-only', round(b['fwd']['ms_per_step'],3))
PY
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/ar_bench.json 2> gpurun_out/ar_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/ar_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, {k:round(v,4) for k,v in b['kernel_ms'].items()})
for k,v in b.get('at_scale',{}).items():
    if isinstance(v,dict): print(k, round(v['ms_per_step'],3), {kk:round(vv,3) for kk,vv in v['kernel_ms'].items()}, 'fwd', round(v.get('fwd_ms_per_step',0),3))
PY
