#!/bin/bash
# chunk-ring K1s: basins per CTA (2 / 4 / 8) and the per-step one-warp ring, training forward
mkdir -p gpurun_out
run() {  # tag, env...
tag=$1; shift
env "$@" timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/ae_c4_$tag.json 2> gpurun_out/ae_c4_$tag.err
python - <<PY
import json
for ln in open('gpurun_out/ae_c4_$tag.json'):
    c=json.loads(ln); print('c4 $tag',round(c['ms_per_step'],3),round(c['fwd_ms_per_step'],3),{kk: round(v,3) for kk,v in c['kernel_ms'].items()},c['checks']['prefix_bit_exact'])
PY
for B in 2500 4000 8000 22500; do
env "$@" timeout 600 python bench.py --workload shard --basins $B --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/ae_b${B}_$tag.json 2> gpurun_out/ae_b${B}_$tag.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/ae_b${B}_$tag.json'))
    print('hbv B=$B $tag ms',round(b['ms_per_step'],3),{kk: round(v,3) for kk,v in b['kernel_ms'].items()}, 'fwd-only', round(b['fwd']['ms_per_step'],3), b['run_info']['ckpt_interval'])
except Exception as e: print('B=$B',e)
PY
done
}
run bpb2 HBV_B200_LEAN_BPB=2
run bpb4 HBV_B200_LEAN_BPB=4
run bpb8 HBV_B200_LEAN_BPB=8
run step HBV_B200_LEAN_SMALL=400000
