#!/bin/bash
mkdir -p gpurun_out
for mode in dma kernel; do
HBV_B200_BLOCK_COPY=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-at-scale --no-cpu-baseline > gpurun_out/j_bench_$mode.json 2> gpurun_out/j_bench_$mode.err
python - <<PY
import json
d=json.load(open('gpurun_out/j_bench_$mode.json'))
print('$mode', 'e2e', {k: d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step','serial_ms_per_step','host_gradient_equals_dense_device_gradient','pcie_GBps')})
PY
done
timeout 600 python -m pytest tests/test_hostio_gpu.py -m gpu -q 2>&1 | tail -2
