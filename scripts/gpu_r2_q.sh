#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lean_gpu.py tests/test_long_hourly_gpu.py tests/test_parity_gpu.py -m gpu -q > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/q_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/q_tests.log | head -30
for k in 1 2 4; do
timeout 600 python bench.py --workload shard --ckpt $k --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/q_shard_k$k.json 2> gpurun_out/q_shard_k$k.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/q_shard_k$k.json'))
    print('shard K=$k ms',b['ms_per_step'],{kk: round(v,3) for kk,v in b['kernel_ms'].items()}, 'fwd-only', b['fwd']['ms_per_step'])
except Exception as e: print('shard K=$k',e)
PY
tail -2 gpurun_out/q_shard_k$k.err
done
