#!/usr/bin/env python
"""Kernel times of hbv_2 (split form, dynamic [parBETA, parK0, parBETAET]) at a throughput-regime
size, for A/B runs of the input path (HBV_B200_DENSE=0/1):  python scripts/ab_hbv2.py [B] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydrodl2_b200 as hydrodl2  # noqa: E402
from hydrodl2_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 22500
T = int(sys.argv[2]) if len(sys.argv) > 2 else 730
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
x = torch.rand(T, B, 3, generator=g, device=dev) * torch.tensor([8.0, 30.0, 4.0], device=dev) - torch.tensor([0.0, 10.0, 0.0], device=dev)
p0 = torch.rand(T, B, 48, generator=g, device=dev).requires_grad_(True)
p1 = torch.rand(B, 13 * 16 + 2, generator=g, device=dev).requires_grad_(True)
xd = {'x_phy': x, 'ac_all': torch.rand(B, generator=g, device=dev) * 5000, 'elev_all': torch.rand(B, generator=g, device=dev) * 3500}
M = hydrodl2.load_model('hbv_2', ver_name='Hbv_2')
m = M({'dynamic_params': {'Hbv_2': ['parBETA', 'parK0', 'parBETAET']}, 'nmul': 16, 'warm_up': 0, 'state_series': False}, device=dev)


def step():
    p0.grad = p1.grad = None
    out = m(xd, [p0, p1])
    out['streamflow'].sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
ops.PROFILE = {}
for _ in range(5):
    step()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
print(f'hbv_2 D3 B={B} T={T} DENSE={os.environ.get("HBV_B200_DENSE", "default")}',
      {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in prof.items()})
