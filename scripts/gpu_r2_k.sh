#!/bin/bash
# 2-GPU: shared-gradient all-reduce captured in the CUDA graph vs eager after the replay
mkdir -p gpurun_out
for ga in 1 0; do
HBV_BENCH_GRAPH_ALLREDUCE=$ga timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-at-scale --no-cpu-baseline > gpurun_out/k_n2_ga$ga.json 2> gpurun_out/k_n2_ga$ga.err
echo "ga=$ga rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/k_n2_ga$ga.json'))
    print('ga=$ga', 'ms', d['ms_per_step'], 'value %.3e'%d['value'], d['run_info']['launch'][:90], 'e2e', d['e2e']['ms_per_step'] if d['e2e'] else None)
except Exception as e: print('ga=$ga', e)
PY
tail -2 gpurun_out/k_n2_ga$ga.err
done
