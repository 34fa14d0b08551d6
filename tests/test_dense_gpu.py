"""GPU parity of the TMA-staged kernels (hbv_dense.cu: K1d / K2d).

The golden fixtures are small and run through the cp.async-ring kernels, so these cases use
basin counts large enough for the dense path (> 2,368 basins at nmul 16; HBV_B200_DENSE=2 takes
it wherever the shapes allow, the default additionally asks for a full-GPU grid) and compare it with
  * the CPU oracle (fp32 restatement of the reference, pinned by tests/test_oracle_golden.py) on
    the same seeded inputs: 1e-5 fluxes/states, 1e-4 parameter gradients (max-norm relative);
  * K1/K2 on the same inputs (HBV_B200_DENSE=0): the step arithmetic is the same code, so the
    results must agree to fp32 round-off of the compiler's contraction choices (5e-6).
Basin counts cover aligned runs (B % 4 == 0), runs whose 16 B phase changes every step
(B % 4 == 2) and odd B (forward dense, adjoint falls back to K2), each with a partial last CTA.
"""

import pytest
import torch

from hydrodl2_b200 import _cabi
from conftest import RTOL_FLUX, RTOL_GRAD, assert_close, assert_grad_close, check_split_slice_vs_oracle

pytestmark = pytest.mark.gpu

NMUL = 16
# two compilations of the same step arithmetic differ by FMA-contraction choices; the largest
# relative difference sits on the excess flux (a cancellation, SM1 - FC): 2e-6 measured
XTOL = 5e-6
D14 = ['parBETA', 'parFC', 'parK0', 'parK1', 'parK2', 'parLP', 'parPERC', 'parUZL', 'parTT',
       'parCFMAX', 'parCFR', 'parCWH', 'parBETAET', 'parC']
D3 = ['parBETA', 'parK0', 'parBETAET']


def _launches():
    from hydrodl2_b200 import _cabi
    return _cabi.launch_count()


def _run_11p(x, p, dev, dense, monkeypatch, cot=None, ckpt=0):
    import hydrodl2_b200 as hydrodl2
    _cabi.set_option('dense', int('2' if dense else '0'))   # 2 = wherever the shapes allow
    M = hydrodl2.load_model('hbv_1_1p', ver_name='Hbv_1_1p')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv_1_1p': D14}, 'nmul': NMUL, 'ckpt_interval': ckpt}, device=dev)
    pg = p.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev)}, pg)
    loss = out['streamflow'].sum() if cot is None else sum((out[k] * c.to(dev)).sum() for k, c in cot.items())
    loss.backward()
    torch.cuda.synchronize()
    return out, pg.grad, m


@pytest.mark.parametrize('B', [2500, 2502, 2501])
def test_dense_hbv_1_1p_vs_oracle(B, monkeypatch):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T = 37
    x = O.synthetic_forcing(T, B, seed=11)
    p = torch.randn(T, B, 14 * NMUL + 2, generator=torch.Generator().manual_seed(12))
    pc = p.clone().requires_grad_(True)
    ref, ref_states = O.forward_packed('hbv_1_1p', x, pc, nmul=NMUL, warm_up=0, dynamic_params=D14)
    ref['streamflow'].sum().backward()

    out, grad, m = _run_11p(x, p, dev, True, monkeypatch)
    for k, v in ref.items():
        assert_close(out[k], v, RTOL_FLUX, f'dense B={B}:{k}')
    assert_grad_close(grad, pc.grad, f'dense B={B}:grad', NMUL)
    for name, s, r in zip(m.state_names, m.get_states(), ref_states):
        assert_close(s, r, RTOL_FLUX, f'dense B={B}:state {name}')

    out0, grad0, _ = _run_11p(x, p, dev, False, monkeypatch)
    for k in ref:
        assert_close(out[k], out0[k], XTOL, f'dense vs K1 B={B}:{k}')
    assert_close(grad, grad0, XTOL, f'dense vs K2 B={B}:grad')


def test_dense_hbv_1_1p_all_series_cotangent(monkeypatch):
    """Upstream gradient on every flux series (not the prefetched streamflow-only case)."""
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B = 21, 2440
    x = O.synthetic_forcing(T, B, seed=13)
    p = torch.randn(T, B, 14 * NMUL + 2, generator=torch.Generator().manual_seed(14))
    out_probe, _, _ = _run_11p(x, p, dev, False, monkeypatch)
    g = torch.Generator().manual_seed(15)
    cot = {k: torch.randn(v.shape, generator=g) for k, v in out_probe.items() if v.dim() == 3}
    out1, grad1, _ = _run_11p(x, p, dev, True, monkeypatch, cot)
    out0, grad0, _ = _run_11p(x, p, dev, False, monkeypatch, cot)
    assert_close(grad1, grad0, XTOL, 'dense vs K2: all-series cotangent grad')
    # K = 16 request: the dense forward writes the sparse checkpoints, the adjoint is K2
    out2, grad2, _ = _run_11p(x, p, dev, True, monkeypatch, cot, ckpt=16)
    assert_close(grad2, grad0, XTOL, 'dense fwd + K2 (K=16): grad')


def test_dense_is_taken(monkeypatch):
    """The dense kernels are the ones that run at this size: same launch count, different
    kernels is not observable from here, so check the library's own dispatch report."""
    from hydrodl2_b200 import _cabi
    lib = _cabi.load()
    if not hasattr(lib, 'hbv_b200_dense_launches'):
        pytest.skip('library built without the dispatch counter')
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B = 9, 2500
    x = O.synthetic_forcing(T, B, seed=1)
    p = torch.randn(T, B, 14 * NMUL + 2, generator=torch.Generator().manual_seed(2))
    n0 = lib.hbv_b200_dense_launches()
    _run_11p(x, p, dev, True, monkeypatch)
    assert lib.hbv_b200_dense_launches() - n0 == 2       # K1d + K2d
    n0 = lib.hbv_b200_dense_launches()
    _run_11p(x, p, dev, False, monkeypatch)
    assert lib.hbv_b200_dense_launches() - n0 == 0


def _run_split(model, cls, x_dict, params, dev, dense, monkeypatch, **cfg):
    import hydrodl2_b200 as hydrodl2
    _cabi.set_option('dense', int('2' if dense else '0'))   # 2 = wherever the shapes allow
    M = hydrodl2.load_model(model, ver_name=cls)
    m = M({'dynamic_params': {cls: D3}, 'nmul': NMUL, **cfg}, device=dev)
    ps = [q.detach().clone().requires_grad_(True) for q in params]
    out = m(x_dict, ps)
    key = 'streamflow' if 'streamflow' in out else 'Qs'
    out[key].sum().backward()
    torch.cuda.synchronize()
    return out, [q.grad for q in ps]


@pytest.mark.parametrize('B', [2500, 2501])
def test_dense_hbv_2_matches_k1_k2(B, monkeypatch):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T = 26
    g = torch.Generator().manual_seed(21)
    x = O.synthetic_forcing(T, B, seed=22).to(dev)
    p0 = torch.rand(T, B, 3 * NMUL, generator=g).to(dev)
    p1 = torch.rand(B, 13 * NMUL + 2, generator=g).to(dev)
    xd = {'x_phy': x, 'ac_all': (torch.rand(B, generator=g) * 5000).to(dev),
          'elev_all': (torch.rand(B, generator=g) * 3500).to(dev)}
    cfg = {'warm_up': 0}
    out1, g1 = _run_split('hbv_2', 'Hbv_2', xd, [p0, p1], dev, True, monkeypatch, **cfg)
    out0, g0 = _run_split('hbv_2', 'Hbv_2', xd, [p0, p1], dev, False, monkeypatch, **cfg)
    for k in out0:
        assert_close(out1[k], out0[k], XTOL, f'hbv_2 dense vs K1 B={B}:{k}')
    for a, b, n in zip(g1, g0, ('dyn', 'static')):
        assert_close(a, b, XTOL, f'hbv_2 dense vs K2 B={B}:grad {n}')
    check_split_slice_vs_oracle('hbv_2', xd, p0, p1, D3, out1, g1, 'streamflow', f'hbv_2 dense B={B}')


def test_dense_hbv_2_hourly_matches_k1_k2(monkeypatch):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B = 50, 2520
    g = torch.Generator().manual_seed(31)
    x = (O.synthetic_forcing(T, B, seed=32) / 24.0).to(dev)
    p0 = torch.rand(T, B, 3 * NMUL, generator=g).to(dev)
    p1 = torch.rand(B, 16 * NMUL, generator=g).to(dev)
    xd = {'x_phy': x, 'ac_all': (torch.rand(B, generator=g) * 5000).to(dev),
          'elev_all': (torch.rand(B, generator=g) * 3500).to(dev)}
    cfg = {'routing': False, 'state_series': False}

    def run(dense):
        import hydrodl2_b200 as hydrodl2
        _cabi.set_option('dense', int('2' if dense else '0'))   # 2 = wherever the shapes allow
        M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
        m = M({'dynamic_params': {'Hbv_2_hourly': D3}, 'nmul': NMUL, **cfg}, device=dev)
        m.use_distr_routing = False
        ps = [q.detach().clone().requires_grad_(True) for q in (p0, p1)]
        out = m(xd, ps)
        out['Qs'].sum().backward()
        torch.cuda.synchronize()
        return out, [q.grad for q in ps]

    out1, g1 = run(True)
    out0, g0 = run(False)
    assert_close(out1['Qs'], out0['Qs'], XTOL, 'hourly dense vs K1: Qs')
    for a, b, n in zip(g1, g0, ('dyn', 'static')):
        assert_close(a, b, XTOL, f'hourly dense vs K2: grad {n}')
    check_split_slice_vs_oracle('hbv_2_hourly', xd, p0, p1, D3, out1, g1, 'Qs', 'hourly dense')
