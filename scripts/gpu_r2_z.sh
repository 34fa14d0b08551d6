#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/z_n$N.json 2> gpurun_out/z_n$N.err ) 2> gpurun_out/z_n$N.time
echo "rc=$?"; cat gpurun_out/z_n$N.time
python - <<PY
import json
d=json.load(open('gpurun_out/z_n$N.json'))
print('N=$N c2 ms/step', d['ms_per_step'], 'value %.4e'%d['value'], d['run_info']['shared_grad_allreduce'])
print('e2e', d['e2e']['ms_per_step'], 'value %.3e'%d['e2e']['value'])
for n, a in d['at_scale'].items():
    if 'error' in a: print(n, a); continue
    print(n, 'ms %.3f fwd-only %.3f value %.3e' % (a['ms_per_step'], a['fwd_ms_per_step'], a['value']), a.get('scaling'))
PY
HBV_BENCH_ONESHOT=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 50 --warmup 5 --no-at-scale --no-cpu-baseline > gpurun_out/z_n${N}_nccl.json 2> gpurun_out/z_n${N}_nccl.err
python - <<PY
import json
d=json.load(open('gpurun_out/z_n${N}_nccl.json'))
print('N=$N NCCL ms/step', d['ms_per_step'], 'value %.4e'%d['value'], 'e2e', d['e2e']['ms_per_step'])
PY
