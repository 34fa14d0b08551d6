#!/bin/bash
# full suite on the K3^T ring / pad-row routing build; K3 forward basins per CTA
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/aj_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/aj_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/aj_tests.log | head
for bpb in 2 4 8; do
HBV_B200_ADJ_BPB=$bpb timeout 600 python scripts/bench_configs.py c5 --steps 5 > gpurun_out/aj_c5_bpb$bpb.json 2> gpurun_out/aj_c5_bpb$bpb.err
python - <<PY
import json
for ln in open('gpurun_out/aj_c5_bpb$bpb.json'):
    c=json.loads(ln); print('c5 bpb$bpb', round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks'])
PY
done
