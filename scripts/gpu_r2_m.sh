#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/m_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/m_tests.log | head -20
for duo in 1 0; do
HBV_B200_DUO=$duo timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/m_bench_duo$duo.json 2> gpurun_out/m_bench_duo$duo.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/m_bench_duo$duo.json'))
    print('duo=$duo c2 ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],{k: round(v,4) for k,v in b['kernel_ms'].items()}, 'e2e', b['e2e']['ms_per_step'])
except Exception as e: print('c2',e)
PY
done
