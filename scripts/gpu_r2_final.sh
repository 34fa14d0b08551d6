#!/bin/bash
# round 2 final evidence: GPU suite, smoke, parity report, K3 ncu capture, compute-sanitizer on the new kernels
mkdir -p gpurun_out /tmp/ncu
( time timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1 ) 2> gpurun_out/final_tests.time
echo "tests rc=$?" >> gpurun_out/final_tests.log
grep -E "FAILED|passed|failed|Error:" gpurun_out/final_tests.log | head; cat gpurun_out/final_tests.time | head -2
python __graft_entry__.py smoke 2>&1 | tail -1
python scripts/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/final_parity.err; tail -3 gpurun_out/final_parity.err
timeout 600 ncu --set full --clock-control none -k regex:hbv_adj_.*_kernel -s 4 -c 2 -f -o /tmp/ncu/r02_c5 python scripts/bench_configs.py c5 --steps 1 > gpurun_out/final_ncu_c5.log 2>&1
ncu -i /tmp/ncu/r02_c5.ncu-rep --page raw --csv > gpurun_out/r02_c5_raw.csv 2>/dev/null
# memcheck + racecheck of the round-2 kernels on small shapes
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import torch
import hydrodl2_b200 as hydrodl2
from hydrodl2_b200 import _cabi
from hydrodl2_b200.hostio import PipelinedSteps
dev = torch.device('cuda:0')
T, B, nmul, warm = 23, 7, 16, 5
g = torch.Generator().manual_seed(1)
x = torch.rand(T, B, 3, generator=g).to(dev) * 5
p = torch.randn(T, B, 13 * nmul + 2, generator=g).to(dev)
for ck in (0, 2, 4):
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': nmul, 'ckpt_interval': ck}, device=dev)
    pg = p.clone().requires_grad_(True)
    m({'x_phy': x}, pg)['streamflow'].sum().backward()
M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
m = M({'dynamic_params': {'Hbv_2_hourly': ['parBETA', 'parK0', 'parBETAET']}, 'nmul': nmul, 'routing': True, 'state_series': False}, device=dev)
p0 = torch.rand(T, B, 48, generator=g).to(dev).requires_grad_(True)
p1 = torch.rand(B, 16 * nmul + 2, generator=g).to(dev).requires_grad_(True)
topo = torch.zeros(2, B); topo[0, :4] = 1; topo[1, 3:] = 1
p2 = torch.rand(int(topo.sum()), 3, generator=g).to(dev).requires_grad_(True)
out = m({'x_phy': x / 24, 'ac_all': torch.rand(B, generator=g).to(dev) * 5000, 'elev_all': torch.rand(B, generator=g).to(dev) * 3500,
         'outlet_topo': topo.to(dev), 'areas': torch.rand(B, generator=g).to(dev) + 1}, [p0, p1, p2])
out['streamflow'].sum().backward()
lib = _cabi.load()
st = torch.cuda.current_stream(dev).cuda_stream
buf = torch.ones(100003, device=dev)
_cabi.check(lib.hbv_b200_fill_zero(buf[1:].data_ptr(), 100000 * 4, 0, st), 'fill')
nfl = int(lib.hbv_b200_allreduce_buffer_floats(1, 33)); cb = torch.zeros(nfl, device=dev)
ptrs = torch.tensor([cb.data_ptr()], dtype=torch.int64, device=dev); v = torch.ones(33, device=dev)
_cabi.check(lib.hbv_b200_oneshot_allreduce(ptrs.data_ptr(), 0, 1, v.data_ptr(), v.data_ptr(), 33, st), 'ar')
torch.cuda.synchronize()
print('sanitizer workload done, pipe launches', lib.hbv_b200_pipe_launches(), 'lean', lib.hbv_b200_lean_launches())
PY
for tool in memcheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/final_san_$tool.log 2>&1
echo "$tool rc=$?"; tail -3 gpurun_out/final_san_$tool.log
done
