#!/bin/bash
# (1) host staging auto-pick; (2) stored-state layout experiment at C4: plane layout / stores kept in L2 / warp-major
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
for v in base exp1 exp2; do
lib=$PWD/hydrodl2_b200/lib/libhbv_b200_$v.so; [ $v = base ] && lib=$PWD/hydrodl2_b200/lib/libhbv_b200.so
HBV_B200_LIB=$lib timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ah_c4_$v.json 2> gpurun_out/ah_c4_$v.err
python - <<PY
import json
for ln in open('gpurun_out/ah_c4_$v.json'):
    c=json.loads(ln); print('c4 $v',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks'])
PY
HBV_B200_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lean_kernel --csv --log-file gpurun_out/ah_c4_${v}_launches.csv python scripts/bench_configs.py c4 --steps 1 > /dev/null 2>&1
grep lean_kernel gpurun_out/ah_c4_${v}_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,140- | tail -4
done
