#!/bin/bash
# round 2, call D: full GPU suite, C2 bench (thin fill), ncu of the C4 recurrence kernels (pipe vs lean)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.log
tail -6 gpurun_out/d_tests.log
for tf in 1 0; do
HBV_B200_THIN_FILL=$tf timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-at-scale > gpurun_out/d_bench_c2_tf$tf.json 2> gpurun_out/d_bench_c2_tf$tf.err
done
for pipe in 1 0; do
HBV_B200_PIPE=$pipe timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_.*_kernel -s 4 -c 2 -f -o gpurun_out/prof_c4_pipe$pipe python scripts/bench_configs.py c4 --steps 1 > gpurun_out/ncu_c4_pipe$pipe.log 2>&1
tail -2 gpurun_out/ncu_c4_pipe$pipe.log
done
python - <<'PY'
import json
for tf in (1,0):
    try:
        b=json.load(open(f'gpurun_out/d_bench_c2_tf{tf}.json'))
        print('c2 thinfill',tf,'ms',b['ms_per_step'],'eager',b['run_info']['eager_ms_per_step'],b['kernel_ms'])
    except Exception as e: print('c2',e)
PY
