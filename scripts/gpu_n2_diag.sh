#!/bin/bash
# 2-GPU: bench line (graph + eager), torchrun as the driver launches it
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-at-scale 2>gpurun_out/n2.err | tail -1 > gpurun_out/bench_n2.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n2.json')); print('N=2', d['config']['launch'], 'step', d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], 'fwd', d['fwd']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" || tail -20 gpurun_out/n2.err
