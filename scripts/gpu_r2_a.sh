#!/bin/bash
# round 2, call A: parity of the stage-pipelined kernels + A/B timing against the round-1 lean kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lean_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
for pipe in 1 0; do
  HBV_B200_PIPE=$pipe timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/a_bench_c2_pipe$pipe.json 2> gpurun_out/a_bench_c2_pipe$pipe.err
  HBV_B200_PIPE=$pipe timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/a_c4_pipe$pipe.json 2> gpurun_out/a_c4_pipe$pipe.err
done
tail -5 gpurun_out/a_tests.log
python - <<'PY'
import json
for pipe in (1,0):
    try:
        b=json.load(open(f'gpurun_out/a_bench_c2_pipe{pipe}.json'))
        print('c2 pipe',pipe,'ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],b['kernel_ms'],'fwd',b['fwd']['kernel_ms'])
    except Exception as e: print('c2',pipe,e)
    try:
        for ln in open(f'gpurun_out/a_c4_pipe{pipe}.json'):
            c=json.loads(ln); print('c4 pipe',pipe,c['ms_per_step'],c['fwd_ms_per_step'],c['kernel_ms'],c['checks'])
    except Exception as e: print('c4',pipe,e)
PY
