#!/bin/bash
# same-box A/B of two builds of the library on the at-scale workloads (REG kernels)
for rep in 1 2; do for lib in libhbv_old.so libhbv_b200.so; do for wl in shard c3; do
HBV_B200_LIB=$PWD/hydrodl2_b200/lib/$lib python bench.py --workload $wl --ckpt 16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $wl', {k: round(v,3) for k,v in d['kernel_ms'].items()})"
done; done; done
