#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_lean_gpu.py tests/test_forcing_grad_gpu.py -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-at-scale 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, 'fwd-only', round(d['fwd']['ms_per_step'],3), d['fwd']['kernel_ms'])
"
timeout 300 python scripts/bench_configs.py c4 --steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('c4', round(d['ms_per_step'],2), round(d['fwd_ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()}, d['checks'])
    except Exception as e: print(l[:200])
"
