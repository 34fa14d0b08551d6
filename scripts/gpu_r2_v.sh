#!/bin/bash
# N-GPU run of the default bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/v_n$N.json 2> gpurun_out/v_n$N.err ) 2> gpurun_out/v_n$N.time
echo "rc=$?"; cat gpurun_out/v_n$N.time; tail -3 gpurun_out/v_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/v_n$N.json'))
print('N=$N c2 ms/step', d['ms_per_step'], 'value %.3e'%d['value'], d['run_info']['launch'][:60])
print('e2e', d['e2e']['ms_per_step'], 'value %.3e'%d['e2e']['value'])
for n, a in d['at_scale'].items():
    if 'error' in a: print(n, a); continue
    print(n, 'ms %.3f fwd-only %.3f value %.3e' % (a['ms_per_step'], a['fwd_ms_per_step'], a['value']), a.get('scaling'))
PY
