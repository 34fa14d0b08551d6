#!/bin/bash
# register-window pair routing + chunk-ring K1s: full suite, C4 timing, launch list, ncu source of both forward forms
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ad_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ad_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ad_tests.log | head
timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ad_c4.json 2> gpurun_out/ad_c4.err
python - <<PY
import json
for ln in open('gpurun_out/ad_c4.json'):
    c=json.loads(ln); print('c4',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ad_c4_launches.csv python scripts/bench_configs.py c4 --steps 1 > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hbv_fwd_lean_kernel -s 2 -c 2 -f -o /tmp/ncu/ad_c4 python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ad_ncu_c4.log 2>&1
ncu -i /tmp/ncu/ad_c4.ncu-rep --page raw --csv > gpurun_out/ad_c4_raw.csv 2>/dev/null
ncu -i /tmp/ncu/ad_c4.ncu-rep --page source --csv --print-source sass > gpurun_out/ad_c4_source.csv 2>/dev/null
ls -la gpurun_out/ad_* | head -20
