"""GPU parity at BASELINE.json's full sizes.

* configs[1] (C2: `hbv`, 531 basins x (365 warm-up + 730) days, dynamic [parBETA, parBETAET]) is
  small enough for the CPU oracle: all series, states and the parameter gradient are compared
  directly (1e-5 / 1e-4).
* the north-star per-GPU shards (22,500 basins x 730 days; `hbv` D2 and configs[2] `hbv_1_1p` with
  all 14 parameters dynamic) are checked through size-independent properties:
    - a 48-basin slice of the full run == the CPU oracle on that slice (basins are independent,
      which is also what multi-GPU sharding rests on), fluxes and gradients;
    - permuting the basins permutes the outputs bit for bit;
    - water balance: sum(P) - sum(AET) - sum(Q) == change in storage, per basin, to 1e-4 of the
      precipitation total (the step only moves water between stores; `hbv.py:428-492`).
"""

import math

import pytest
import torch

from conftest import RTOL_FLUX, RTOL_GRAD, assert_close, assert_grad_close

pytestmark = pytest.mark.gpu

NMUL = 16
D2 = ['parBETA', 'parBETAET']
D14 = ['parBETA', 'parFC', 'parK0', 'parK1', 'parK2', 'parLP', 'parPERC', 'parUZL', 'parTT',
       'parCFMAX', 'parCFR', 'parCWH', 'parBETAET', 'parC']


def _model(name, cls, dyn, warm_up, dev):
    import hydrodl2_b200 as hydrodl2
    M = hydrodl2.load_model(name, ver_name=cls)
    return M({'warm_up': warm_up, 'dynamic_params': {cls: dyn}, 'nmul': NMUL}, device=dev)


def test_c2_full_size_vs_oracle():
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, warm = 365 + 730, 531, 365
    x = O.synthetic_forcing(T, B, seed=20261017)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(20261018))
    pc = p.clone().requires_grad_(True)
    ref, ref_states = O.forward_packed('hbv', x, pc, nmul=NMUL, warm_up=warm, dynamic_params=D2)
    ref['streamflow'].sum().backward()
    m = _model('hbv', 'Hbv', D2, warm, dev)
    pg = p.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev)}, pg)
    out['streamflow'].sum().backward()
    for k, v in ref.items():
        assert_close(out[k], v, RTOL_FLUX, f'C2 full size: {k}')
    for name, s, r in zip(m.state_names, m.get_states(), ref_states):
        assert_close(s, r, RTOL_FLUX, f'C2 full size: state {name}')
    assert_grad_close(pg.grad, pc.grad, 'C2 full size: grad', NMUL)


def _device_inputs(T, B, ncol, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    d = torch.arange(T, dtype=torch.float32, device=dev).view(T, 1)
    ob = torch.rand(1, B, generator=g, device=dev) * 16 - 8
    season = torch.sin(2 * math.pi * (d - 110) / 365)
    tmean = 5 + 12 * season + ob + 4 * torch.randn(T, B, generator=g, device=dev)
    prcp = 5 * torch.relu(torch.randn(T, B, generator=g, device=dev))
    pet = torch.relu(2 + 2 * season) + 0.5 * torch.rand(T, B, generator=g, device=dev)
    x = torch.stack([prcp, tmean, pet], dim=-1).contiguous()
    p = torch.randn(T, B, ncol, generator=g, device=dev)
    return x, p


@pytest.mark.parametrize('name,cls,npar,dyn', [('hbv', 'Hbv', 13, D2), ('hbv_1_1p', 'Hbv_1_1p', 14, D14)])
def test_shard_full_size_properties(name, cls, npar, dyn):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    # slice of 512 basins (SURVEY §8 d6 asks for 2,048: the CPU oracle's backward takes ~200 s on such
    # a slice, 512 keeps the suite within the driver's limit; basins are independent, so the slice
    # size changes the statistics of the comparison, not what is compared)
    T, B, nb = 730, 22500, 512
    x, p = _device_inputs(T, B, npar * NMUL + 2, dev, seed=5)
    m = _model(name, cls, dyn, 0, dev)
    pg = p.clone().requires_grad_(True)
    out = m({'x_phy': x}, pg)
    out['streamflow'].sum().backward()
    torch.cuda.synchronize()
    for k, v in out.items():
        assert torch.isfinite(v).all(), k
    assert torch.isfinite(pg.grad).all()

    # (1) a slice of the full run == the oracle on that slice (taken from the middle, so the
    #     basins sit in CTAs with neighbours on both sides)
    lo = 11111
    xs, ps = x[:, lo:lo + nb].cpu(), p[:, lo:lo + nb].cpu()
    pc = ps.clone().requires_grad_(True)
    ref, ref_states = O.forward_packed(name, xs, pc, nmul=NMUL, warm_up=0, dynamic_params=dyn)
    ref['streamflow'].sum().backward()
    for k, v in ref.items():
        got = out[k][lo:lo + nb] if k == 'BFI' else out[k][:, lo:lo + nb]
        assert_close(got, v, RTOL_FLUX, f'{name} shard slice: {k}')
    assert_grad_close(pg.grad[:, lo:lo + nb], pc.grad, f'{name} shard slice: grad', NMUL)
    for sname, s, r in zip(m.state_names, m.get_states(), ref_states):
        assert_close(s[lo:lo + nb], r, RTOL_FLUX, f'{name} shard slice: state {sname}')

    # (2) water balance per basin (mean over the components): P - AET - Q = d(storage)
    P = x[:, :, 0].sum(0)
    aet = out['AET_hydro'][:, :, 0].sum(0)
    q = out['streamflow_no_rout'][:, :, 0].sum(0)
    storage_end = torch.stack([s.mean(-1) for s in m.get_states()]).sum(0)
    storage_0 = 5 * 0.001
    resid = (P - aet - q) - (storage_end - storage_0)
    assert (resid.abs() <= 1e-4 * P.abs().max()).all(), f'water balance residual {resid.abs().max().item():.3e}'

    # (3) permuting the basins permutes the outputs bit for bit
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(9)).to(dev)
    m2 = _model(name, cls, dyn, 0, dev)
    with torch.no_grad():
        out_p = m2({'x_phy': x[:, perm].contiguous()}, p[:, perm].contiguous())
        out_0 = m2({'x_phy': x}, p)
    for k in ('streamflow', 'AET_hydro', 'SWE'):
        assert torch.equal(out_p[k], out_0[k][:, perm]), f'{name}: permutation changed {k}'
    assert torch.equal(out_p['BFI'], out_0['BFI'][perm])
