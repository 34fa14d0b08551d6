#!/bin/bash
# stored-state layout 0 / 1 at C4 and on the shard (forced through the option)
mkdir -p gpurun_out
for lay in 0 1; do
HBV_B200_CKPT_LAYOUT=$lay timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/av_c4.json 2> gpurun_out/av_c4.err
python - <<PY
import json
for ln in open('gpurun_out/av_c4.json'):
    c=json.loads(ln); print('c4 layout=$lay',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['prefix_bit_exact'])
PY
HBV_B200_CKPT_LAYOUT=$lay timeout 600 python bench.py --workload shard --steps 10 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/av_shard.json 2> gpurun_out/av_shard.err
python - <<PY
import json
b=json.load(open('gpurun_out/av_shard.json'))
print('shard layout=$lay ms',round(b['ms_per_step'],3),{kk: round(v,3) for kk,v in b['kernel_ms'].items()})
PY
done
