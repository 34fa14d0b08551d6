#!/bin/bash
# dense (TMA-staged) kernels: parity, then A/B against K1/K2 on the c3 workload
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense_gpu.py -x -q 2>&1 | tail -15
for cfg in "0 0 0" "1 0 0" "1 4 3" "1 8 4"; do
set -- $cfg
HBV_B200_DENSE=$1 HBV_B200_DENSE_NS=$2 HBV_B200_DENSE_NS_BWD=$3 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dense_$1_$2_$3.json 2>gpurun_out/dense.err || tail -5 gpurun_out/dense.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/dense_$1_$2_$3.json'))
    print('c3 dense=$1 ns=$2 nsb=$3', 'ms %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
except Exception as e: print('failed', e)
PY
done
