"""Host-side cost of one c2 training step (cProfile over 300 steps, GPU box)."""
import cProfile, pstats, sys, os, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import hydrodl2_b200 as hydrodl2
dev = torch.device('cuda:0')
wl = bench.WORKLOADS['c2']
x, p = bench.make_inputs(wl, wl['B'], 1)
x, p = x.to(dev), p.to(dev).requires_grad_(True)
M = hydrodl2.load_model('hbv', ver_name='Hbv')
m = M(bench.model_config(wl), device=dev)
def step():
    p.grad = None
    out = m({'x_phy': x}, p)
    out['streamflow'].sum().backward()
for _ in range(20): step()
torch.cuda.synchronize()
# pure host time: enqueue only, sync at the end
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3*(t1-t0)/300:.3f} ms/step; with final sync {1e3*(t2-t0)/300:.3f} ms/step')
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
