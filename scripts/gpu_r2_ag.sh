#!/bin/bash
# host staging: copy-engine vs SM-driven block copies picked by timing; e2e on this box with each forced mode
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hostio_gpu.py -m gpu -q > gpurun_out/ag_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ag_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/ag_tests.log | head
for mode in auto dma kernel; do
HBV_B200_BLOCK_COPY=$mode timeout 600 python bench.py --no-at-scale --no-cpu-baseline --steps 50 > gpurun_out/ag_bench_$mode.json 2> gpurun_out/ag_bench_$mode.err
python - <<PY
import json
b=json.load(open('gpurun_out/ag_bench_$mode.json')); e=b['e2e']
print('$mode', round(b['ms_per_step'],4), 'e2e', round(e['ms_per_step'],3), e['pcie_GBps'], e.get('block_copy'), e['host_gradient_equals_dense_device_gradient'])
PY
done
nproc; lscpu | grep -E "Model name|Socket|NUMA" | head -5; nvidia-smi topo -m 2>/dev/null | head -6
