#!/usr/bin/env python
"""Host-side profile of the eager C2 training step (cProfile over 300 steps): where the Python /
ctypes / ATen time of the public-API path goes when it is not replayed from a CUDA graph."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydrodl2_b200 as hydrodl2  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    T, B, nmul, warm = 1095, 531, 16, 365
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': nmul}, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand(T, B, 3, generator=g, device=dev) * 5
    p = torch.randn(T, B, 13 * nmul + 2, generator=g, device=dev).requires_grad_(True)

    def step():
        p.grad = None
        out = m({'x_phy': x}, p)
        out['streamflow'].sum().backward()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        step()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f'eager: host issue {t_issue / 300 * 1e3:.3f} ms/step, wall {t_all / 300 * 1e3:.3f} ms/step')
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(45)
    st.sort_stats('tottime').print_stats(30)


if __name__ == '__main__':
    main()
