#!/bin/bash
# round 2, call F: C2 bench + C4 timing after the carve-out fix; new golden cases on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_lean_gpu.py -m gpu -q -x > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/f_tests.log
tail -5 gpurun_out/f_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err
timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/f_c4.json 2> gpurun_out/f_c4.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/f_bench_c2.json'))
    print('c2 ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],b['kernel_ms'])
except Exception as e: print('c2',e)
try:
    for ln in open('gpurun_out/f_c4.json'):
        c=json.loads(ln); print('c4',c['ms_per_step'],c['fwd_ms_per_step'],c['kernel_ms'],c['checks'])
except Exception as e: print('c4',e)
PY
