#!/bin/bash
# round 2, call B: parity of the re-pipelined kernels, timing, and an ncu capture of the three C2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lean_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "not test_lean_packed_vs_oracle or True" > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b_tests.log
HBV_B200_PIPE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/b_bench_c2.json 2> gpurun_out/b_bench_c2.err
HBV_B200_PIPE=1 timeout 600 python scripts/bench_configs.py c4 --steps 3 > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hbv_.*pipe_kernel -s 9 -c 3 -f -o gpurun_out/prof_c2_pipe python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/ncu_c2_pipe.log 2>&1
tail -2 gpurun_out/ncu_c2_pipe.log
tail -8 gpurun_out/b_tests.log
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/b_bench_c2.json'))
    print('c2 ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],b['kernel_ms'],'fwd',b['fwd']['kernel_ms'])
except Exception as e: print('c2',e)
try:
    for ln in open('gpurun_out/b_c4.json'):
        c=json.loads(ln); print('c4',c['ms_per_step'],c['fwd_ms_per_step'],c['kernel_ms'],c['checks'])
except Exception as e: print('c4',e)
PY
