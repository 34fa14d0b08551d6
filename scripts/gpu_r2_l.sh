#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/l_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/l_tests.log | head -30
timeout 600 python scripts/bench_configs.py c5 --steps 5 > gpurun_out/l_c5.json 2> gpurun_out/l_c5.err
python - <<'PY'
import json
for ln in open('gpurun_out/l_c5.json'):
    c=json.loads(ln); print('c5',c['ms_per_step'],c['fwd_ms_per_step'],c['kernel_ms'],c['checks'])
PY
