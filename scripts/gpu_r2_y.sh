#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_allreduce_gpu.py -m gpu -q 2>&1 | tail -3
for os_ in 1 0; do
HBV_BENCH_ONESHOT=$os_ timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$os_ bench.py --gpus $N --steps 50 --warmup 5 --no-at-scale --no-cpu-baseline > gpurun_out/y_n${N}_os$os_.json 2> gpurun_out/y_n${N}_os$os_.err
echo "oneshot=$os_ rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/y_n${N}_os$os_.json'))
    print('N=$N oneshot=$os_ ms', d['ms_per_step'], 'value %.4e'%d['value'], d['run_info']['shared_grad_allreduce'], '|', d['run_info']['launch'][:70], '| e2e', d['e2e']['ms_per_step'])
except Exception as e: print('N=$N oneshot=$os_', e)
PY
grep -v "UserWarning\|return func\|^\*\*\*\|OMP_NUM" gpurun_out/y_n${N}_os$os_.err | tail -4
done
