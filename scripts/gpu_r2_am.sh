#!/bin/bash
# e2e with mixed block-copy modes (upload, download) and copy-kernel grid sizes
mkdir -p gpurun_out
for mode in dma,dma dma,kernel kernel,dma kernel,kernel; do
for cb in 1 2 4; do
HBV_B200_COPY_BLOCKS=$cb HBV_B200_BLOCK_COPY=$mode timeout 600 python bench.py --no-at-scale --no-cpu-baseline --steps 50 > gpurun_out/am_bench.json 2> gpurun_out/am_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/am_bench.json')); e=b['e2e']
print('$mode blocks/SM $cb', round(b['ms_per_step'],4), 'e2e', round(e['ms_per_step'],3), {k: round(v,1) for k,v in e['pcie_GBps'].items()}, e.get('block_copy'), e['host_gradient_equals_dense_device_gradient'])
PY
[ $mode = dma,dma ] && break
done
done
