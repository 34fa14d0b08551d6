#!/bin/bash
# chunk-ring K1s (cp.async, 128-thread CTAs): parity + timing against the register form
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_lean_gpu.py tests/test_parity_gpu.py tests/test_long_hourly_gpu.py tests/test_fullsize_gpu.py tests/test_dense_gpu.py -m gpu -q -x > gpurun_out/ac_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ac_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ac_tests.log | head
for deep in -1 0; do
HBV_B200_LEAN_DEEP=$deep timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ac_c4_deep$deep.json 2> gpurun_out/ac_c4_deep$deep.err
python - <<PY
import json
for ln in open('gpurun_out/ac_c4_deep$deep.json'):
    c=json.loads(ln); print('c4 deep=$deep',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['prefix_bit_exact'])
PY
for B in 2500 4000 8000 22500; do
HBV_B200_LEAN_DEEP=$deep timeout 600 python bench.py --workload shard --basins $B --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/ac_b${B}_deep$deep.json 2> gpurun_out/ac_b${B}_deep$deep.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/ac_b${B}_deep$deep.json'))
    print('hbv B=$B deep=$deep ms',round(b['ms_per_step'],3),{kk: round(v,3) for kk,v in b['kernel_ms'].items()}, 'fwd-only', round(b['fwd']['ms_per_step'],3), b['run_info']['ckpt_interval'])
except Exception as e: print('B=$B',e)
PY
done
done
