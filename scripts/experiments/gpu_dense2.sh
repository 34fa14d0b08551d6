#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense_gpu.py tests/test_hostio_gpu.py -x -q 2>&1 | tail -5
for bpb in 8 4 2; do
HBV_B200_DENSE_BPB=$bpb timeout 300 python -m pytest tests/test_dense_gpu.py -x -q 2>&1 | tail -1
HBV_B200_DENSE_BPB=$bpb timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dense_bpb$bpb.json 2>gpurun_out/dense.err || tail -5 gpurun_out/dense.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/dense_bpb$bpb.json'))
    print('c3 bwd bpb=$bpb', 'ms %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
except Exception as e: print('failed', e)
PY
done
python bench.py --workload shard --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/shard_b.json 2>gpurun_out/dense.err || tail -5 gpurun_out/dense.err
python - <<PY
import json
d=json.load(open('gpurun_out/shard_b.json'))
print('shard', 'ms %.3f fwd-only %.3f' % (d['ms_per_step'], d['fwd']['ms_per_step']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
PY
