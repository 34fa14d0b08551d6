#!/usr/bin/env python
"""CPU baseline of BASELINE configs 3-5 on bounded slices, host cores only:

    python scripts/cpu_slices.py [--basins 256] > profiles/r02_cpu_slices.json

The unmodified reference (baseline/_ref) where it can run, timed on a slice of the basins — its
cost is linear in the basin count — and, for the backward of config 4, on a shorter time axis
(its autograd is superlinear in T); config 5 (`hbv_adj`) cannot be imported in the
reference (encrypted batch_jacobian.pye), so the float32 PyTorch restatement is timed instead and
labelled so.  Every entry states exactly what was run.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    nb = int(sys.argv[sys.argv.index('--basins') + 1]) if '--basins' in sys.argv else 256
    print(json.dumps(run(nb), indent=1))


def run(nb=256, nb_hourly=64, nb_adj=64):
    """-> dict of timings; importable (bench.py's cpu_baseline leg runs it on smaller slices)."""
    import bench
    from oracle import hbv_oracle as O
    ref = bench.load_reference()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = 'reference' if ref is not None else 'port'
    out = {'cores': cores, 'kind': kind, 'slice_basins': nb}
    NMUL = 16

    # ---- config 3: hbv_1_1p, all 14 parameters dynamic, 730 days, fwd and fwd+bwd
    D14 = bench.D14
    x = O.synthetic_forcing(730, nb, seed=31)
    p = torch.randn(730, nb, 14 * NMUL + 2, generator=torch.Generator().manual_seed(32))
    if ref is not None:
        M = ref.load_model('hbv_1_1p', ver_name='Hbv_1_1p')
        m = M({'warm_up': 0, 'dynamic_params': {'Hbv_1_1p': D14}, 'nmul': NMUL}, device=torch.device('cpu'))
        run = lambda pp: m({'x_phy': x}, pp)                                     # noqa: E731
    else:
        run = lambda pp: O.forward_packed('hbv_1_1p', x, pp, nmul=NMUL, warm_up=0, dynamic_params=D14)[0]   # noqa: E731
    t0 = time.perf_counter()
    with torch.no_grad():
        run(p)
    tf = time.perf_counter() - t0
    pr = p.clone().requires_grad_(True)
    t0 = time.perf_counter()
    run(pr)['streamflow'].sum().backward()
    tb = time.perf_counter() - t0
    out['c3'] = {'what': f'hbv_1_1p D14, {nb} basins x 730 days ({kind})', 'fwd_s': tf, 'fwd_bwd_s': tb,
                 'fwd_basin_steps_per_s': nb * 730 / tf, 'fwd_bwd_basin_steps_per_s': nb * 730 / tb}

    # ---- config 4: hbv_2_hourly D3; forward on the full 17,520 steps, fwd+bwd on 2,160 steps
    dyn = ['parBETA', 'parK0', 'parBETAET']
    for T, leg in ((17520, 'fwd'), (2160, 'fwd_bwd')):
        nbh = min(nb, nb_hourly)
        g = torch.Generator().manual_seed(41)
        xh = O.synthetic_forcing(T, nbh, seed=42, hourly=True)
        p0 = torch.rand(T, nbh, 3 * NMUL, generator=g)
        p1 = torch.rand(nbh, 16 * NMUL, generator=g)
        xd = {'x_phy': xh, 'ac_all': torch.rand(nbh, generator=g) * 5000, 'elev_all': torch.rand(nbh, generator=g) * 3500}
        if ref is not None:
            M = ref.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
            mh = M({'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': NMUL, 'routing': False}, device=torch.device('cpu'))
            mh.use_distr_routing = False
            xd2 = dict(xd, outlet_topo=torch.eye(nbh), areas=torch.ones(nbh))
            runh = lambda a, b: mh(xd2, [a, b, torch.rand(nbh, 3)])              # noqa: E731
        else:
            runh = lambda a, b: O.forward_split('hbv_2_hourly', xd, [a, b], nmul=NMUL, dynamic_params=dyn,   # noqa: E731
                                                use_distr_routing=False)[0]
        t0 = time.perf_counter()
        if leg == 'fwd':
            with torch.no_grad():
                runh(p0, p1)
        else:
            a, b = p0.clone().requires_grad_(True), p1.clone().requires_grad_(True)
            runh(a, b)['Qs'].sum().backward()
        dt = time.perf_counter() - t0
        out[f'c4_{leg}'] = {'what': f'hbv_2_hourly D3, {nbh} units x {T} hourly steps, {leg} ({kind})', 'seconds': dt,
                            'basin_steps_per_s': nbh * T / dt}

    # ---- config 5: hbv_adj — the reference cannot run; the restatement is timed
    from oracle import hbv_adj_oracle as OA
    nb5 = min(nb, nb_adj)
    x5 = O.synthetic_forcing(730, nb5, seed=51)
    p5 = torch.randn(730, nb5, 13 * NMUL + 2, generator=torch.Generator().manual_seed(52)).requires_grad_(True)
    t0 = time.perf_counter()
    try:
        res = OA.forward_adj(x5, p5, nmul=NMUL, warm_up=0, dynamic_params=['parBETA', 'parBETAET'])
        res['flow_sim'].sum().backward()
        dt = time.perf_counter() - t0
        out['c5'] = {'what': f'hbv_adj float32 PyTorch restatement (oracle/hbv_adj_oracle.py; the reference is not '
                             f'runnable), {nb5} basins x 730 days, fwd+bwd', 'seconds': dt,
                     'basin_steps_per_s': nb5 * 730 / dt}
    except Exception as exc:
        out['c5'] = {'error': f'{type(exc).__name__}: {exc}'}
    return out


if __name__ == '__main__':
    main()
