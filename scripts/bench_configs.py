#!/usr/bin/env python
"""Run the BASELINE.json configurations that are not the bench line at their per-GPU sizes and
print one JSON object per configuration (GPU box):

    python scripts/bench_configs.py [c4] [c5] [--steps 3]

  c4  hbv_2_hourly, 2,500 units/GPU (20k / 8) x 17,520 hourly steps, dynamic
      [parBETA, parK0, parBETAET], distributed pair routing (gage g drains 40 units), fwd+bwd,
      long-sequence checkpointed adjoint, state series not materialised
  c5  hbv_adj implicit scheme, 10,000 basins x 730 days (the 1-GPU size of configs[4]), fwd+bwd

Beyond timing, each run checks size-independent properties at full size (SURVEY.md §8 c):
finite outputs, non-negative flows, water balance of the run (c4: sum(P) - sum(Qs) bounded by
storage + ET), and that a 64-basin prefix run alone reproduces the same prefix of the full run
bit-for-bit (basins are independent — the property the multi-GPU sharding rests on).
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

NMUL = 16


def forcing(T, B, dev, seed, hourly=False):
    g = torch.Generator(device=dev).manual_seed(seed)
    d = torch.arange(T, dtype=torch.float32, device=dev).view(T, 1) / (24.0 if hourly else 1.0)
    ob = torch.rand(1, B, generator=g, device=dev) * 16 - 8
    season = torch.sin(2 * math.pi * (d - 110) / 365)
    tmean = 5 + 12 * season + ob + 4 * torch.randn(T, B, generator=g, device=dev)
    prcp = 5 * torch.relu(torch.randn(T, B, generator=g, device=dev))
    pet = torch.relu(2 + 2 * season) + 0.5 * torch.rand(T, B, generator=g, device=dev)
    x = torch.stack([prcp, tmean, pet], dim=-1).contiguous()
    return x / 24.0 if hourly else x


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_c4(steps, B=2500, dev=None, seed_offset=0):
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200 import ops
    dev = torch.device('cuda:0') if dev is None else dev
    T, per_gage = 17520, 40
    dyn = ['parBETA', 'parK0', 'parBETAET']
    g = torch.Generator(device=dev).manual_seed(4 + seed_offset)
    x = forcing(T, B, dev, 40 + seed_offset, hourly=True)
    p0 = torch.rand(T, B, 3 * NMUL, generator=g, device=dev).requires_grad_(True)
    p1 = torch.rand(B, 16 * NMUL, generator=g, device=dev).requires_grad_(True)
    n_gage = (B + per_gage - 1) // per_gage
    topo = torch.zeros(n_gage, B, device=dev)
    for gi in range(n_gage):
        topo[gi, gi * per_gage:(gi + 1) * per_gage] = 1.0
    areas = torch.rand(B, generator=g, device=dev) * 99 + 1
    p2 = torch.rand(int(topo.sum().item()), 3, generator=g, device=dev).requires_grad_(True)
    xd = {'x_phy': x, 'ac_all': torch.rand(B, generator=g, device=dev) * 5000,
          'elev_all': torch.rand(B, generator=g, device=dev) * 3500, 'outlet_topo': topo, 'areas': areas}
    M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
    m = M({'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': NMUL, 'routing': False,
           'state_series': False}, device=dev)

    def step():
        for p in (p0, p1, p2):
            p.grad = None
        out = m(xd, [p0, p1, p2])
        out['streamflow'].sum().backward()
        return out

    def fwd():
        with torch.no_grad():
            return m(xd, [p0, p1, p2])

    ops.PROFILE = {}
    ms = timed(step, steps)
    prof, ops.PROFILE = ops.PROFILE, None
    # (the last `steps` calls of every kernel group: the warm-up calls carry the lazy module load)
    kms = {k: sum(a.elapsed_time(b) for a, b in v[-steps:]) / len(v[-steps:]) for k, v in prof.items()}
    ms_f = timed(fwd, steps)
    out = step()
    torch.cuda.synchronize()
    qs, sf = out['Qs'], out['streamflow']
    checks = {
        'finite': bool(torch.isfinite(qs).all() and torch.isfinite(sf).all()
                       and torch.isfinite(p0.grad).all() and torch.isfinite(p1.grad).all()),
        # PERC = min(SUZ, pc)/dt*dt can overshoot SUZ by an ulp (reference arithmetic): allow -1e-6
        'nonnegative_flow': bool((qs >= -1e-6 * qs.max()).all() and (sf >= -1e-6 * sf.max()).all()),
        'runoff_le_precip_plus_lateral': bool(qs.sum() <= x[..., 0].sum() * 1.5 + 1e3),
    }
    # basin independence: a 64-unit prefix alone == the prefix of the full run (bit-exact)
    xs = {k: (v[:, :64].contiguous() if k == 'x_phy' else v[:64]) for k, v in xd.items()
          if k in ('x_phy', 'ac_all', 'elev_all')}
    m2 = M({'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': NMUL, 'routing': False, 'state_series': False},
           device=dev)
    m2.use_distr_routing = False
    # (same mode as the full run — gradients on — so the same kernel instantiation computes both)
    nsub = min(64, B)
    from hydrodl2_b200 import _cabi
    with _cabi.option('pipe', 0):      # (the same kernel family as the full grid: K1s / K2s)
        sub = m2(xs, [p0[:, :nsub].detach().contiguous().requires_grad_(True), p1[:nsub].detach().contiguous()])
    checks['prefix_bit_exact'] = bool(torch.equal(sub['Qs'].detach(), qs[:, :64].detach()))
    n_dyn = 3
    bf = 4 * (3 + n_dyn * NMUL + 1) + 5 * NMUL * 4 / 16
    bb = 4 * (3 + 2 * n_dyn * NMUL + 1) + 5 * NMUL * 4 / 16
    return {
        'config': f'c4: hbv_2_hourly fwd+bwd, {B} units/GPU x {T} hourly steps, nmul 16, dynamic {dyn}, '
                  f'{n_gage} gages x {per_gage} units pair routing, state series off',
        'ms_per_step': ms, 'fwd_ms_per_step': ms_f, 'basin_timesteps_per_s': B * T / (ms * 1e-3),
        'fwd_basin_timesteps_per_s': B * T / (ms_f * 1e-3), 'kernel_ms': kms,
        'hbm_GBps_fwd_kernel': bf * B * T / (kms['hbv_fwd'] * 1e-3) / 1e9,
        'hbm_GBps_bwd_kernel': bb * B * T / (kms['hbv_bwd'] * 1e-3) / 1e9,
        'algorithmic_bytes_per_basin_step': {'fwd': bf, 'bwd': bb}, 'checks': checks,
        'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9,
        'units_per_gpu': B * T, 'fwd_kernel': 'hbv_fwd', 'bwd_kernel': 'hbv_bwd',
        # ours: every state stored (K = 1: 320 B written by the forward, read by the adjoint) and the
        # gradient's three dynamic blocks written; survey: SURVEY.md §8 (d4) worked figures
        'bytes': {'hbv_fwd': {'ours': 4.0 * (3 + n_dyn * NMUL + 1) + 320, 'survey': 208.0},
                  'hbv_bwd': {'ours': 4.0 * (3 + 2 * n_dyn * NMUL + 1) + 320, 'survey': 420.0}},
    }


def run_c5(steps, B=10000, dev=None, seed_offset=0):
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200 import ops
    dev = torch.device('cuda:0') if dev is None else dev
    T = 730
    dyn = ['parBETA', 'parBETAET']
    x = forcing(T, B, dev, 50 + seed_offset)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator(device=dev).manual_seed(5 + seed_offset),
                    device=dev).requires_grad_(True)
    M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
    m = M({'warm_up': 0, 'dynamic_params': {'HbvAdj': dyn}, 'nmul': NMUL}, device=dev)

    def step():
        p.grad = None
        out = m({'x_phy': x}, p)
        out['flow_sim'].sum().backward()
        return out

    def fwd():
        with torch.no_grad():
            return m({'x_phy': x}, p)

    ops.PROFILE = {}
    ms = timed(step, steps)
    prof, ops.PROFILE = ops.PROFILE, None
    # (the last `steps` calls of every kernel group: the warm-up calls carry the lazy module load)
    kms = {k: sum(a.elapsed_time(b) for a, b in v[-steps:]) / len(v[-steps:]) for k, v in prof.items()}
    ms_f = timed(fwd, steps)
    out = step()
    torch.cuda.synchronize()
    q = out['flow_sim']
    stats = m.newton_stats.cpu().tolist()
    m2 = M({'warm_up': 0, 'dynamic_params': {'HbvAdj': dyn}, 'nmul': NMUL}, device=dev)
    with torch.no_grad():
        sub = m2({'x_phy': x[:, :64].contiguous()}, p[:, :64].contiguous())
    checks = {
        'finite': bool(torch.isfinite(q).all() and torch.isfinite(p.grad).all()),
        'nonnegative_flow': bool((q >= 0).all()),
        'newton_converged_everywhere': stats[1] == 0, 'newton_max_updates_used': stats[0],
        'prefix_bit_exact': bool(torch.equal(sub['flow_sim'], q[:, :64])),
    }
    return {
        'config': f'c5: hbv_adj implicit scheme fwd+bwd, {B} basins x {T} days, nmul 16, dynamic {dyn}',
        'ms_per_step': ms, 'fwd_ms_per_step': ms_f, 'basin_timesteps_per_s': B * T / (ms * 1e-3),
        'fwd_basin_timesteps_per_s': B * T / (ms_f * 1e-3), 'kernel_ms': kms, 'checks': checks,
        'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9,
        'units_per_gpu': B * T, 'fwd_kernel': 'hbv_adj_fwd', 'bwd_kernel': 'hbv_adj_bwd',
        # forcings + 2 dynamic blocks + Qsim + every converged state; the survey
        # gives no worked figure for the implicit scheme
        'bytes': {'hbv_adj_fwd': {'ours': 4.0 * (3 + 32 + 1) + 320, 'survey': None},
                  'hbv_adj_bwd': {'ours': 4.0 * (3 + 64 + 1) + 320, 'survey': None}},
    }


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    steps = 3
    if '--steps' in sys.argv:
        steps = int(sys.argv[sys.argv.index('--steps') + 1])
        args = [a for a in args if a != str(steps)]
    which = args or ['c4', 'c5']
    for w in which:
        t0 = time.time()
        res = {'c4': run_c4, 'c5': run_c5}[w](steps)
        res['wall_s'] = time.time() - t0
        print(json.dumps(res), flush=True)
        from hydrodl2_b200 import ops
        ops.release_grad_planes()      # (idle cached gradient planes of the finished workload)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == '__main__':
    main()
