#!/bin/bash
# round 2 closing evidence on the shipped build: suite, smoke, parity report, launch list of the bench
# command, ncu --set full of the C2 / shard / C4 / C5 kernels (exported to CSV on the box), sanitizer
mkdir -p gpurun_out /tmp/ncu
rm -f gpurun_out/*.ncu-rep
( time timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1 ) 2> gpurun_out/final_tests.time
echo "tests rc=$?" >> gpurun_out/final_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/final_tests.log | head; head -2 gpurun_out/final_tests.time
python __graft_entry__.py smoke 2>&1 | tail -1
python scripts/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/final_parity.err; tail -2 gpurun_out/final_parity.err
timeout 900 python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-at-scale --no-cpu-baseline > gpurun_out/t_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_.*pipe_kernel -s 9 -c 3 -f -o /tmp/ncu/r02_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale --no-graph > gpurun_out/t_ncu_c2.log 2>&1
ncu -i /tmp/ncu/r02_c2.ncu-rep --page raw --csv > gpurun_out/r02_c2_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:hbv_.*lean_kernel -s 6 -c 3 -f -o /tmp/ncu/r02_shard python bench.py --workload shard --steps 1 --warmup 3 --no-cpu-baseline --no-at-scale > gpurun_out/t_ncu_shard.log 2>&1
ncu -i /tmp/ncu/r02_shard.ncu-rep --page raw --csv > gpurun_out/r02_shard_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:hbv_.*_lean_kernel -s 4 -c 2 -f -o /tmp/ncu/r02_c4 python scripts/bench_configs.py c4 --steps 1 > gpurun_out/t_ncu_c4.log 2>&1
ncu -i /tmp/ncu/r02_c4.ncu-rep --page raw --csv > gpurun_out/r02_c4_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:hbv_adj_.*_kernel -s 4 -c 2 -f -o /tmp/ncu/r02_c5 python scripts/bench_configs.py c5 --steps 1 > gpurun_out/t_ncu_c5.log 2>&1
ncu -i /tmp/ncu/r02_c5.ncu-rep --page raw --csv > gpurun_out/r02_c5_raw.csv 2>/dev/null
# memcheck + racecheck of the round-2 kernels on small shapes
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import torch
import hydrodl2_b200 as hydrodl2
from hydrodl2_b200 import _cabi
dev = torch.device('cuda:0')
nmul = 16
g = torch.Generator().manual_seed(1)
# one-warp / stage-pipelined kernels (small grid) at K = 1, 2, 4
T, B, warm = 23, 7, 5
x = torch.rand(T, B, 3, generator=g).to(dev) * 5
p = torch.randn(T, B, 13 * nmul + 2, generator=g).to(dev)
for ck in (0, 2, 4):
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': nmul, 'ckpt_interval': ck}, device=dev)
    pg = p.clone().requires_grad_(True)
    m({'x_phy': x}, pg)['streamflow'].sum().backward()
# 128-thread chunk-ring forward + ring adjoint (grid above the small-grid threshold), K = 1 and 4
T2, B2 = 11, 2403
x2 = torch.rand(T2, B2, 3, generator=g).to(dev) * 5
p2 = torch.randn(T2, B2, 13 * nmul + 2, generator=g).to(dev)
for ck in (1, 4):
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 2, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': nmul, 'ckpt_interval': ck}, device=dev)
    pg = p2.clone().requires_grad_(True)
    m({'x_phy': x2}, pg)['streamflow'].sum().backward()
# hourly model with pair routing (register-window conv, tap-parallel gamma kernels, segmented sums)
M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
m = M({'dynamic_params': {'Hbv_2_hourly': ['parBETA', 'parK0', 'parBETAET']}, 'nmul': nmul, 'routing': True, 'state_series': False}, device=dev)
T3 = 150
x3 = torch.rand(T3, B, 3, generator=g).to(dev) * 5
p0 = torch.rand(T3, B, 48, generator=g).to(dev).requires_grad_(True)
p1 = torch.rand(B, 16 * nmul + 2, generator=g).to(dev).requires_grad_(True)
topo = torch.zeros(2, B); topo[0, :4] = 1; topo[1, 3:] = 1
pr = torch.rand(int(topo.sum()), 3, generator=g).to(dev).requires_grad_(True)
out = m({'x_phy': x3 / 24, 'ac_all': torch.rand(B, generator=g).to(dev) * 5000, 'elev_all': torch.rand(B, generator=g).to(dev) * 3500,
         'outlet_topo': topo.to(dev), 'areas': torch.rand(B, generator=g).to(dev) + 1}, [p0, p1, pr])
out['streamflow'].sum().backward()
# implicit scheme: one-warp forward, ring adjoint with the fused zero fill
M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
m = M({'warm_up': 3, 'dynamic_params': {'HbvAdj': ['parBETA', 'parBETAET']}, 'nmul': nmul}, device=dev)
pg = p.clone().requires_grad_(True)
m({'x_phy': x}, pg)['flow_sim'].sum().backward()
lib = _cabi.load()
st = torch.cuda.current_stream(dev).cuda_stream
buf = torch.ones(100003, device=dev)
_cabi.check(lib.hbv_b200_fill_zero(buf[1:].data_ptr(), 100000 * 4, 0, st), 'fill')
nfl = int(lib.hbv_b200_allreduce_buffer_floats(1, 33)); cb = torch.zeros(nfl, device=dev)
ptrs = torch.tensor([cb.data_ptr()], dtype=torch.int64, device=dev); v = torch.ones(33, device=dev)
_cabi.check(lib.hbv_b200_oneshot_allreduce(ptrs.data_ptr(), 0, 1, v.data_ptr(), v.data_ptr(), 33, st), 'ar')
torch.cuda.synchronize()
print('sanitizer workload done, launches', lib.hbv_b200_launch_count(), 'pipe', lib.hbv_b200_pipe_launches(), 'lean', lib.hbv_b200_lean_launches())
PY
for tool in memcheck racecheck; do
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/final_san_$tool.log 2>&1
echo "$tool rc=$?"; tail -3 gpurun_out/final_san_$tool.log
done
ls -la gpurun_out/r02_* gpurun_out/final_* | head -30
