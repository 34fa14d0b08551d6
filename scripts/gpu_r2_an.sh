#!/bin/bash
# N-GPU bench: one-shot all-reduce overlapped with the next step vs in line
N=${1:-2}
mkdir -p gpurun_out
for ov in ${OVS:-1 0}; do
HBV_BENCH_OVERLAP_ALLREDUCE=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus $N --no-at-scale --no-cpu-baseline > gpurun_out/an_n${N}_ov$ov.json 2> gpurun_out/an_n${N}_ov$ov.err
echo "rc=$?"; tail -2 gpurun_out/an_n${N}_ov$ov.err
python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/an_n${N}_ov$ov.json') if l.startswith('{')][-1])
    print('N=$N overlap=$ov', round(b['ms_per_step'],4), 'value', '%.4g'%b['value'], 'e2e', round(b['e2e']['ms_per_step'],3), b['run_info']['shared_grad_allreduce'])
except Exception as e: print('parse', e)
PY
done
