"""GPU parity of the gradients w.r.t. forcings and w.r.t. `muwts`: the adjoint
kernel's by-products against PyTorch autograd over the CPU oracle (the reference's arithmetic).
Tolerance: max-norm relative 1e-4, the parameter-gradient bar."""

import pytest
import torch

from conftest import RTOL_FLUX, RTOL_GRAD, assert_close, assert_grad_close

pytestmark = pytest.mark.gpu


def _cot(ref, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(v.shape, generator=g) for k, v in ref.items()}


@pytest.mark.parametrize('model,cls,dyn,npar,nmul', [
    ('hbv', 'Hbv', ['parBETA', 'parBETAET'], 13, 16),
    ('hbv_1_1p', 'Hbv_1_1p', ['parBETA', 'parK0', 'parBETAET'], 14, 4),
    ('hbv', 'Hbv', [], 12, 3),          # nmul not a power of two: atomic reduction path
])
def test_forcing_gradient_packed(model, cls, dyn, npar, nmul):
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B, warm = 120, 9, 30
    x = O.synthetic_forcing(T, B, seed=21)
    p = torch.randn(T, B, npar * nmul + 2, generator=torch.Generator().manual_seed(22))
    xc, pc = x.clone().requires_grad_(True), p.clone().requires_grad_(True)
    ref, _ = O.forward_packed(model, xc, pc, nmul=nmul, warm_up=warm, dynamic_params=dyn)
    cot = _cot(ref, 23)
    sum((ref[k] * cot[k]).sum() for k in ref).backward()

    M = hydrodl2.load_model(model, ver_name=cls)
    m = M({'warm_up': warm, 'dynamic_params': {cls: dyn}, 'nmul': nmul}, device=dev)
    xg, pg = x.to(dev).requires_grad_(True), p.to(dev).requires_grad_(True)
    out = m({'x_phy': xg}, pg)
    sum((out[k] * cot[k].to(dev)).sum() for k in ref).backward()
    for k in ref:
        assert_close(out[k], ref[k], RTOL_FLUX, k)
    assert_grad_close(pg.grad, pc.grad, 'grad parameters', nmul)
    assert xg.grad.shape == x.shape
    assert float(xg.grad[:warm].abs().max()) == 0.0      # warm-up runs under no_grad (hbv.py:328)
    for c, name in enumerate(('prcp', 'tmean', 'pet')):
        assert_close(xg.grad[..., c], xc.grad[..., c], RTOL_GRAD, f'grad x_phy[{name}]')


def test_forcing_gradient_only_forcing_requires_grad():
    """Variational precipitation DA use: parameters fixed, d(loss)/d(prcp) wanted."""
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B, nmul = 90, 7, 16
    x = O.synthetic_forcing(T, B, seed=31)
    p = torch.randn(T, B, 12 * nmul + 2, generator=torch.Generator().manual_seed(32))
    xc = x.clone().requires_grad_(True)
    ref, _ = O.forward_packed('hbv', xc, p, nmul=nmul, warm_up=0, dynamic_params=[])
    ref['streamflow'].sum().backward()
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv': []}, 'nmul': nmul}, device=dev)
    xg = x.to(dev).requires_grad_(True)
    out = m({'x_phy': xg}, p.to(dev))
    out['streamflow'].sum().backward()
    assert_close(xg.grad, xc.grad, RTOL_GRAD, 'grad x_phy')


def test_forcing_gradient_hourly_split():
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B, nmul = 96, 6, 4
    dyn = ['parBETA', 'parK0', 'parBETAET']
    g = torch.Generator().manual_seed(41)
    x = O.synthetic_forcing(T, B, seed=42, hourly=True)
    pd = torch.rand(T, B, 3 * nmul, generator=g)
    ps = torch.rand(B, 16 * nmul, generator=g)
    ac, el = torch.rand(B, generator=g) * 5000, torch.rand(B, generator=g) * 3500
    xd = {'x_phy': x.clone().requires_grad_(True), 'ac_all': ac, 'elev_all': el}
    pdc, psc = pd.clone().requires_grad_(True), ps.clone().requires_grad_(True)
    ref, _ = O.forward_split('hbv_2_hourly', xd, (pdc, psc), nmul=nmul, dynamic_params=dyn,
                             routing=False, use_distr_routing=False)
    ref['Qs'].sum().backward()

    M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
    m = M({'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul, 'routing': False}, device=dev)
    m.use_distr_routing = False
    xg = x.to(dev).requires_grad_(True)
    pdg, psg = pd.to(dev).requires_grad_(True), ps.to(dev).requires_grad_(True)
    out = m({'x_phy': xg, 'ac_all': ac.to(dev), 'elev_all': el.to(dev)}, (pdg, psg))
    out['Qs'].sum().backward()
    assert_close(out['Qs'], ref['Qs'], RTOL_FLUX, 'Qs')
    assert_close(pdg.grad, pdc.grad, RTOL_GRAD, 'grad dyn')
    assert_close(xg.grad, xd['x_phy'].grad, RTOL_GRAD, 'grad x_phy')


@pytest.mark.parametrize('time_varying', [False, True])
def test_muwts_gradient(time_varying):
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B, nmul = 80, 5, 16
    x = O.synthetic_forcing(T, B, seed=51)
    p = torch.randn(T, B, 12 * nmul + 2, generator=torch.Generator().manual_seed(52))
    mu = torch.softmax(torch.randn((T if time_varying else 1), B, nmul,
                                   generator=torch.Generator().manual_seed(53)), dim=-1)
    pc, muc = p.clone().requires_grad_(True), mu.clone().requires_grad_(True)
    ref, _ = O.forward_packed('hbv', x, pc, nmul=nmul, warm_up=0, dynamic_params=[], muwts=muc)
    cot = _cot(ref, 54)
    sum((ref[k] * cot[k]).sum() for k in ref).backward()
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv': []}, 'nmul': nmul}, device=dev)
    pg, mug = p.to(dev).requires_grad_(True), mu.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev), 'muwts': mug}, pg)
    sum((out[k] * cot[k].to(dev)).sum() for k in ref).backward()
    for k in ref:
        assert_close(out[k], ref[k], RTOL_FLUX, k)
    assert_grad_close(pg.grad, pc.grad, 'grad parameters', nmul)
    assert mug.grad.shape == mu.shape
    assert_close(mug.grad, muc.grad, RTOL_GRAD, 'grad muwts')
