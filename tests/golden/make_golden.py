"""Generate golden input/output/gradient fixtures from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What it does: copies ``/root/reference/src/hydrodl2`` to
a scratch directory under /tmp, adds the ``_version.py`` stub that hatch-vcs
would generate (``hydrodl2/__init__.py:11,17`` refuses to import without it),
imports it with ``CI=1`` (licence prompt bypass), runs each model's own
``forward`` on seeded synthetic inputs on CPU and stores inputs, every output
series, the final states and the autograd gradient of a seeded random
cotangent w.r.t. the raw parameters in ``tests/golden/<case>.npz``.

No reference source is copied into the repository; only numbers are.
"""

import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.hbv_oracle import synthetic_forcing  # noqa: E402  (input generator only)


def import_reference():
    scratch = tempfile.mkdtemp(prefix='hydroref_')
    shutil.copytree('/root/reference/src/hydrodl2', os.path.join(scratch, 'hydrodl2'))
    with open(os.path.join(scratch, 'hydrodl2', '_version.py'), 'w') as f:
        f.write("__version__ = '1.0.0+ref'\n")
    os.environ['CI'] = '1'
    sys.path.insert(0, scratch)
    import hydrodl2
    return hydrodl2


def cotangent(out, seed):
    g = torch.Generator().manual_seed(seed)
    loss = 0.0
    cots = {}
    for k in sorted(out.keys()):
        v = out[k]
        if v is None or not v.requires_grad:
            continue
        c = torch.rand(v.shape, generator=g)
        cots[k] = c
        loss = loss + (v * c).sum()
    return loss, cots


def save(name, **arrs):
    if ONLY and name not in ONLY:
        return
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f'{k}/{kk}'] = vv.detach().numpy() if torch.is_tensor(vv) else np.asarray(vv)
        elif torch.is_tensor(v):
            flat[k] = v.detach().numpy()
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **flat)
    print(name, f'{os.path.getsize(path) / 1024:.0f} KiB')


def packed_case(hydrodl2, case, model, cls, dyn, T, B, nmul, warm_up, dy_drop, seed,
                warm_up_states=True):
    M = hydrodl2.load_model(model, ver_name=cls)
    cfg = {'warm_up': warm_up, 'dynamic_params': {cls: dyn}, 'nmul': nmul,
           'dy_drop': dy_drop, 'warm_up_states': warm_up_states}
    m = M(cfg, device=torch.device('cpu'))
    x = synthetic_forcing(T, B, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    p = torch.randn(T, B, m.learnable_param_count, generator=g).requires_grad_(True)
    torch.manual_seed(seed + 2)  # consumed by the reference's bernoulli draws
    out = m({'x_phy': x}, p)
    loss, cots = cotangent(out, seed + 3)
    loss.backward()
    save(case, x_phy=x, parameters=p, grad_parameters=p.grad, out=out, cot=cots,
         states=dict(zip(m.state_names, m.get_states())),
         meta=np.array([T, B, nmul, warm_up, seed + 2]), dy_drop=dy_drop,
         dyn=np.array(dyn, dtype='U16'), model=model,
         warm_up_states=int(warm_up_states))


def split_case(hydrodl2, case, model, cls, dyn, T, B, nmul, dy_drop, seed, routing=False, **extra_cfg):
    M = hydrodl2.load_model(model, ver_name=cls)
    cfg = {'dynamic_params': {cls: dyn}, 'nmul': nmul, 'dy_drop': dy_drop,
           'routing': routing}
    cfg.update(extra_cfg)
    m = M(cfg, device=torch.device('cpu'))
    hourly = model == 'hbv_2_hourly'
    x = synthetic_forcing(T, B, seed=seed, hourly=hourly)
    g = torch.Generator().manual_seed(seed + 1)
    p0 = torch.rand(T, B, m.learnable_param_count1, generator=g).requires_grad_(True)
    p1 = torch.rand(B, m.learnable_param_count2, generator=g).requires_grad_(True)
    xd = {'x_phy': x, 'ac_all': torch.rand(B, generator=g) * 5000,
          'elev_all': torch.rand(B, generator=g) * 3500}
    params = [p0, p1]
    extra = {}
    if hourly:
        ng = 3
        topo = torch.zeros(ng, B)
        for u in range(B):
            topo[u % ng, u] = 1
        topo[0, B - 1] = 1  # one unit drains to two gages
        xd['outlet_topo'] = topo
        xd['areas'] = torch.rand(B, generator=g) * 99 + 1
        p2 = torch.rand(int(topo.sum()), 3, generator=g).requires_grad_(True)
        params.append(p2)
        extra = {'outlet_topo': topo, 'areas': xd['areas']}
    torch.manual_seed(seed + 2)
    out = m(xd, params)
    loss, cots = cotangent(out, seed + 3)
    loss.backward()
    grads = {'p0': p0.grad, 'p1': p1.grad}
    if hourly:
        grads['p2'] = params[2].grad
    save(case, x_phy=x, p0=p0, p1=p1, ac_all=xd['ac_all'], elev_all=xd['elev_all'],
         out=out, cot=cots, grad=grads,
         series=dict(zip(m.state_names, m._state_cache)),
         meta=np.array([T, B, nmul, 0, seed + 2]), dy_drop=dy_drop,
         dyn=np.array(dyn, dtype='U16'), model=model, routing=int(routing),
         warm_up=int(extra_cfg.get('warm_up', 0)), warm_up_states=int(extra_cfg.get('warm_up_states', True)),
         **({'p2': params[2]} if hourly else {}), **extra)


def mts_case(hydrodl2, case, T_low, T_high, B, nmul, seed):
    M = hydrodl2.load_model('hbv_2_mts', ver_name='Hbv_2_mts')
    dyn = ['parBETA', 'parK0', 'parBETAET']
    lo_cfg = {'dynamic_params': {'Hbv_2': dyn}, 'nmul': nmul, 'cache_states': True}
    hi_cfg = {'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul,
              'train_spatial_chunk_size': 10 ** 6, 'simulate_spatial_chunk_size': 10 ** 6,
              'simulate_temporal_chunk_size': 10 ** 6, 'train_warmup': 0}
    m = M(lo_cfg, hi_cfg, device=torch.device('cpu'))
    g = torch.Generator().manual_seed(seed + 1)
    xl = synthetic_forcing(T_low, B, seed=seed)
    xh = synthetic_forcing(T_high, B, seed=seed + 5, hourly=True)
    lo = m.low_freq_model
    hi = m.high_freq_model
    lo_dyn = torch.rand(T_low, B, lo.learnable_param_count1, generator=g).requires_grad_(True)
    lo_sta = torch.rand(B, lo.learnable_param_count2, generator=g).requires_grad_(True)
    hi_dyn = torch.rand(T_high, B, hi.learnable_param_count1, generator=g).requires_grad_(True)
    hi_sta = torch.rand(B, hi.learnable_param_count2, generator=g).requires_grad_(True)
    topo = torch.zeros(2, B)
    topo[0, :B // 2] = 1
    topo[1, B // 2:] = 1
    areas = torch.rand(B, generator=g) * 99 + 1
    hi_distr = torch.rand(int(topo.sum()), 3, generator=g)
    xd = {'x_phy_low_freq': xl, 'x_phy_high_freq': xh, 'ac_all': torch.rand(B, generator=g) * 5000,
          'elev_all': torch.rand(B, generator=g) * 3500, 'outlet_topo': topo, 'areas': areas}
    torch.manual_seed(seed + 2)
    out = m(xd, ([lo_dyn, lo_sta], [hi_dyn, hi_sta, hi_distr]))
    loss, cots = cotangent(out, seed + 3)
    loss.backward()
    zero = torch.zeros(1)
    save(case, x_low=xl, x_high=xh, ac_all=xd['ac_all'], elev_all=xd['elev_all'], outlet_topo=topo,
         areas=areas, lo_dyn=lo_dyn, lo_sta=lo_sta, hi_dyn=hi_dyn, hi_sta=hi_sta, hi_distr=hi_distr,
         out=out, cot=cots,
         grad={'lo_dyn': lo_dyn.grad if lo_dyn.grad is not None else zero,
               'lo_sta': lo_sta.grad, 'hi_dyn': hi_dyn.grad, 'hi_sta': hi_sta.grad},
         meta=np.array([T_low, T_high, B, nmul, seed + 2]), dyn=np.array(dyn, dtype='U16'))


ONLY = set(sys.argv[1:])


def main():
    hydrodl2 = import_reference()
    D2 = ['parBETA', 'parBETAET']
    packed_case(hydrodl2, 'hbv_static', 'hbv', 'Hbv', [], 96, 5, 16, 24, 0.0, 100)
    packed_case(hydrodl2, 'hbv_d2', 'hbv', 'Hbv', D2, 96, 5, 16, 24, 0.0, 200)
    packed_case(hydrodl2, 'hbv_d2_drop_nowarm', 'hbv', 'Hbv', D2, 64, 6, 4, 16, 0.5, 300,
                warm_up_states=False)
    packed_case(hydrodl2, 'hbv_1_1p_d3', 'hbv_1_1p', 'Hbv_1_1p',
                ['parBETA', 'parK0', 'parBETAET'], 96, 5, 16, 24, 0.0, 400)
    all14 = ['parBETA', 'parFC', 'parK0', 'parK1', 'parK2', 'parLP', 'parPERC', 'parUZL',
             'parTT', 'parCFMAX', 'parCFR', 'parCWH', 'parBETAET', 'parC']
    packed_case(hydrodl2, 'hbv_1_1p_d14', 'hbv_1_1p', 'Hbv_1_1p', all14, 64, 4, 16, 0, 0.0, 500)
    split_case(hydrodl2, 'hbv_2_d3', 'hbv_2', 'Hbv_2', ['parBETA', 'parK0', 'parBETAET'],
               96, 5, 16, 0.0, 600)
    split_case(hydrodl2, 'hbv_2_d3_rout', 'hbv_2', 'Hbv_2', ['parBETA', 'parK0', 'parBETAET'],
               64, 5, 4, 0.3, 650, routing=True)
    split_case(hydrodl2, 'hbv_2_hourly_d3', 'hbv_2_hourly', 'Hbv_2_hourly',
               ['parBETA', 'parK0', 'parBETAET'], 120, 6, 16, 0.0, 700)
    mts_case(hydrodl2, 'hbv_2_mts_train', 40, 96, 6, 4, 800)
    # round 2: warm_up > 0 with warm_up_states=False on a split model (the 2.0 models never set
    # `pred_cutoff`, so all T rows come back — hbv_2.py:666-669), and the hourly model's per-unit
    # gamma-UH routing with lenF = 72 (hbv_2_hourly.py:684-705) under the pair routing
    split_case(hydrodl2, 'hbv_2_d3_nowarm', 'hbv_2', 'Hbv_2', ['parBETA', 'parK0', 'parBETAET'],
               48, 5, 16, 0.0, 900, warm_up=8, warm_up_states=False)
    split_case(hydrodl2, 'hbv_2_hourly_rout72', 'hbv_2_hourly', 'Hbv_2_hourly',
               ['parBETA', 'parK0', 'parBETAET'], 150, 6, 16, 0.0, 950, routing=True)


if __name__ == '__main__':
    main()
