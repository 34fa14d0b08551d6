"""GPU parity of the standard-layout kernels (hbv_lean.cu: K1s / K2s).

They serve the throughput regime (large grids), so the cases use > 2,368 basins and compare
  * `hbv` / `hbv_1_1p` with the shipped dynamic set against the CPU oracle on the same seeded
    inputs (1e-5 fluxes / states, 1e-4 gradients), with warm-up, partial last CTA, odd B;
  * every variant against K1 / K2 on the same inputs (HBV_B200_LEAN=0): same step arithmetic,
    agreement to fp32 contraction noise (5e-6);
and check through the library's dispatch counter that the lean kernels are the ones that ran.
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
D2 = ['parBETA', 'parBETAET']
D3 = ['parBETA', 'parK0', 'parBETAET']


def _lean_count():
    from hydrodl2_b200 import _cabi
    return _cabi.load().hbv_b200_lean_launches()


def _run_packed(model, cls, npar, x, p, dev, lean, monkeypatch, warm_up, ckpt=0):
    import hydrodl2_b200 as hydrodl2
    _cabi.set_option('lean', int('1' if lean else '0'))
    M = hydrodl2.load_model(model, ver_name=cls)
    m = M({'warm_up': warm_up, 'dynamic_params': {cls: D2}, 'nmul': NMUL, 'ckpt_interval': ckpt}, device=dev)
    pg = p.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev)}, pg)
    out['streamflow'].sum().backward()
    torch.cuda.synchronize()
    return out, pg.grad, m


@pytest.mark.parametrize('model,cls,npar', [('hbv', 'Hbv', 13), ('hbv_1_1p', 'Hbv_1_1p', 14)])
@pytest.mark.parametrize('B', [2500, 2501])
def test_lean_packed_vs_oracle(model, cls, npar, B, monkeypatch):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, warm = 41, 6
    x = O.synthetic_forcing(T, B, seed=41)
    p = torch.randn(T, B, npar * NMUL + 2, generator=torch.Generator().manual_seed(42))
    pc = p.clone().requires_grad_(True)
    ref, ref_states = O.forward_packed(model, x, pc, nmul=NMUL, warm_up=warm, dynamic_params=D2)
    ref['streamflow'].sum().backward()

    n0 = _lean_count()
    out, grad, m = _run_packed(model, cls, npar, x, p, dev, True, monkeypatch, warm)
    assert _lean_count() - n0 == 3, 'warm-up K1s + K1s + K2s should have run'
    for k, v in ref.items():
        assert_close(out[k], v, RTOL_FLUX, f'lean {model} B={B}:{k}')
    assert_grad_close(grad, pc.grad, f'lean {model} B={B}:grad', NMUL)
    for name, s, r in zip(m.state_names, m.get_states(), ref_states):
        assert_close(s, r, RTOL_FLUX, f'lean {model} B={B}:state {name}')

    n0 = _lean_count()
    out0, grad0, _ = _run_packed(model, cls, npar, x, p, dev, False, monkeypatch, warm)
    assert _lean_count() - n0 == 0
    for k in ref:
        assert_close(out[k], out0[k], XTOL, f'lean vs K1 {model} B={B}:{k}')
    assert_close(grad, grad0, XTOL, f'lean vs K2 {model} B={B}:grad')


@pytest.mark.parametrize('B', [47, 2501])                 # small grid (ring forward) and large (register forward)
@pytest.mark.parametrize('ckpt', [2, 4])
def test_lean_segment_sweep_vs_oracle(ckpt, B, monkeypatch):
    """K1s / K2s with a state stored every 2nd / 4th step (the segment sweep recomputes the missing
    ones): same results as the oracle, partial last segment included (35 run steps)."""
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, warm = 41, 6
    nb = min(B, 64)
    x = O.synthetic_forcing(T, B, seed=43)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(44))
    pc = p[:, :nb].clone().requires_grad_(True)
    ref, ref_states = O.forward_packed('hbv', x[:, :nb], pc, nmul=NMUL, warm_up=warm, dynamic_params=D2)
    ref['streamflow'].sum().backward()
    n0 = _lean_count()
    out, grad, m = _run_packed('hbv', 'Hbv', 13, x, p, dev, True, monkeypatch, warm, ckpt=ckpt)
    assert _lean_count() - n0 == 3, f'K = {ckpt}: warm-up K1s + K1s + K2s should have run'
    for k, v in ref.items():
        got = out[k][:nb] if k == 'BFI' else out[k][:, :nb]
        assert_close(got, v, RTOL_FLUX, f'lean K={ckpt} B={B}:{k}')
    assert_grad_close(grad[:, :nb], pc.grad, f'lean K={ckpt} B={B}:grad', NMUL)
    out1, grad1, _ = _run_packed('hbv', 'Hbv', 13, x, p, dev, True, monkeypatch, warm, ckpt=1)
    assert_close(grad, grad1, XTOL, f'lean K={ckpt} vs K=1 B={B}:grad')


def test_lean_forward_only_no_grad(monkeypatch):
    """Inference (no checkpoints): K1s without the state stores."""
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B = 23, 2440
    x = O.synthetic_forcing(T, B, seed=43)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(44))
    ref, _ = O.forward_packed('hbv', x, p, nmul=NMUL, warm_up=0, dynamic_params=D2)
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv': D2}, 'nmul': NMUL}, device=dev)
    n0 = _lean_count()
    with torch.no_grad():
        out = m({'x_phy': x.to(dev)}, p.to(dev))
    assert _lean_count() - n0 == 1
    for k, v in ref.items():
        assert_close(out[k], v, RTOL_FLUX, f'lean fwd-only:{k}')


def test_lean_not_taken_for_other_cotangents(monkeypatch):
    """A loss on more than the streamflow series goes to K2 and still matches the oracle."""
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B = 19, 2400
    x = O.synthetic_forcing(T, B, seed=45)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(46))
    pc = p.clone().requires_grad_(True)
    ref, _ = O.forward_packed('hbv', x, pc, nmul=NMUL, warm_up=0, dynamic_params=D2)
    (ref['streamflow'].sum() + 0.5 * ref['AET_hydro'].sum()).backward()
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv': D2}, 'nmul': NMUL}, device=dev)
    pg = p.to(dev).requires_grad_(True)
    n0 = _lean_count()
    out = m({'x_phy': x.to(dev)}, pg)
    (out['streamflow'].sum() + 0.5 * out['AET_hydro'].sum()).backward()
    assert _lean_count() - n0 == 1        # K1s forward, K2 adjoint
    assert_grad_close(pg.grad, pc.grad, 'two-series loss: grad', NMUL)


def _split_inputs(B, T, n_static_cols, hourly, seed):
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(seed)
    x = O.synthetic_forcing(T, B, seed=seed + 1)
    if hourly:
        x = x / 24.0
    p0 = torch.rand(T, B, 3 * NMUL, generator=g).to(dev)
    p1 = torch.rand(B, n_static_cols, generator=g).to(dev)
    xd = {'x_phy': x.to(dev), 'ac_all': (torch.rand(B, generator=g) * 5000).to(dev),
          'elev_all': (torch.rand(B, generator=g) * 3500).to(dev)}
    return xd, p0, p1


@pytest.mark.parametrize('B', [2500, 2501])
def test_lean_hbv_2_matches_k1_k2(B, monkeypatch):
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    xd, p0, p1 = _split_inputs(B, 26, 13 * NMUL + 2, False, 51)

    def run(lean):
        _cabi.set_option('lean', int('1' if lean else '0'))
        M = hydrodl2.load_model('hbv_2', ver_name='Hbv_2')
        m = M({'dynamic_params': {'Hbv_2': D3}, 'nmul': NMUL, 'warm_up': 0, 'state_series': False}, device=dev)
        ps = [q.detach().clone().requires_grad_(True) for q in (p0, p1)]
        n0 = _lean_count()
        out = m(xd, ps)
        out['streamflow'].sum().backward()
        torch.cuda.synchronize()
        return out, [q.grad for q in ps], _lean_count() - n0

    out1, g1, n1 = run(True)
    out0, g0, n0 = run(False)
    assert (n1, n0) == (2, 0)
    for k in out0:
        assert_close(out1[k], out0[k], XTOL, f'hbv_2 lean vs K1 B={B}:{k}')
    for a, b, n in zip(g1, g0, ('dyn', 'static')):
        assert_close(a, b, XTOL, f'hbv_2 lean vs K2 B={B}:grad {n}')
    # oracle leg (not only transitively through K1 / K2)
    check_split_slice_vs_oracle('hbv_2', xd, p0, p1, D3, out1, g1, 'streamflow', f'hbv_2 lean B={B}')


def test_lean_hbv_2_hourly_matches_k1_k2(monkeypatch):
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    xd, p0, p1 = _split_inputs(2520, 50, 16 * NMUL, True, 61)

    def run(lean):
        _cabi.set_option('lean', int('1' if lean else '0'))
        M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
        m = M({'dynamic_params': {'Hbv_2_hourly': D3}, 'nmul': NMUL, 'routing': False, 'state_series': False}, device=dev)
        m.use_distr_routing = False
        ps = [q.detach().clone().requires_grad_(True) for q in (p0, p1)]
        n0 = _lean_count()
        out = m(xd, ps)
        out['Qs'].sum().backward()
        torch.cuda.synchronize()
        return out, [q.grad for q in ps], _lean_count() - n0

    out1, g1, n1 = run(True)
    out0, g0, n0 = run(False)
    assert (n1, n0) == (2, 0)
    assert_close(out1['Qs'], out0['Qs'], XTOL, 'hourly lean vs K1: Qs')
    for a, b, n in zip(g1, g0, ('dyn', 'static')):
        assert_close(a, b, XTOL, f'hourly lean vs K2: grad {n}')
    check_split_slice_vs_oracle('hbv_2_hourly', xd, p0, p1, D3, out1, g1, 'Qs', 'hourly lean')


def test_lean_fused_zero_fill_on_poisoned_memory(monkeypatch):
    """K2s writing every element of the gradient rows itself (one-warp form, gdyn_zero_fill) into
    uninitialised memory == the memset path, including the untouched warm-up rows / routing
    columns the host clears."""
    from oracle import hbv_oracle as O
    from hydrodl2_b200 import ops
    dev = torch.device('cuda:0')
    T, B, warm = 33, 2501, 4
    x = O.synthetic_forcing(T, B, seed=71)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(72))
    grads = {}
    prev = ops.FUSED_ZERO_FILL
    try:
        for fused in (False, True):
            ops.FUSED_ZERO_FILL = fused
            torch.cuda.empty_cache()
            junk = torch.empty(T * B * (13 * NMUL + 2) + 4096, device=dev).fill_(float('nan'))
            del junk                                   # poison the block the plane will reuse
            n0 = _lean_count()
            _, g, _ = _run_packed('hbv', 'Hbv', 13, x, p, dev, True, monkeypatch, warm)
            assert _lean_count() - n0 == 3
            grads[fused] = g
    finally:
        ops.FUSED_ZERO_FILL = prev
    assert torch.isfinite(grads[True]).all()
    assert_close(grads[True], grads[False], 1e-7, 'fused zero fill vs memset')
    assert (grads[True][:warm] == 0).all()


@pytest.mark.parametrize('T,B,warm', [(1, 1, 0), (2, 3, 0), (3, 1, 1), (5, 2, 2), (13, 7, 0), (17, 33, 4)])
def test_lean_tiny_shapes(T, B, warm, monkeypatch):
    """Edge sizes of the small-grid (cp.async ring) forms: fewer steps than the ring is deep, a single
    basin, a partial warp, a one-step warm-up."""
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    x = O.synthetic_forcing(T, B, seed=81)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(82))
    pc = p.clone().requires_grad_(True)
    ref, ref_states = O.forward_packed('hbv', x, pc, nmul=NMUL, warm_up=warm, dynamic_params=D2)
    ref['streamflow'].sum().backward()
    n0 = _lean_count()
    out, grad, m = _run_packed('hbv', 'Hbv', 13, x, p, dev, True, monkeypatch, warm)
    assert _lean_count() - n0 == (3 if warm else 2)      # (warm-up K1s +) K1s + K2s
    for k, v in ref.items():
        assert_close(out[k], v, RTOL_FLUX, f'tiny T={T} B={B}:{k}')
    assert_grad_close(grad, pc.grad, f'tiny T={T} B={B}:grad', NMUL)
    for name, s, r in zip(m.state_names, m.get_states(), ref_states):
        assert_close(s, r, RTOL_FLUX, f'tiny T={T} B={B}:state {name}')


@pytest.mark.parametrize('lean,ckpt', [(True, 1), (True, 4), (False, 1), (False, 16)])
@pytest.mark.parametrize('B', [2500, 2501, 1203])
def test_state_store_layouts_give_identical_results(B, lean, ckpt, monkeypatch):
    """hbv_desc_t.ckpt_layout: planes over all lanes (0) or warp-major (1).  Only addresses
    differ, so outputs and gradients are bit-identical — through K1s / K2s (every-state and
    segment sweep; chunk-ring and one-warp forward: 2,500 / 1,203 basins; odd B = a half-filled
    last 32-lane group) and through the generic K1 / K2 (K = 1 ring and K = 16 recompute), which
    are everyone's fallback."""
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, warm = 29, 4
    x = O.synthetic_forcing(T, B, seed=51)
    p = torch.randn(T, B, 13 * NMUL + 2, generator=torch.Generator().manual_seed(52))
    res = {}
    for layout in (0, 1):
        _cabi.set_option('ckpt_layout', layout)
        n0 = _lean_count()
        out, grad, _ = _run_packed('hbv', 'Hbv', 13, x, p, dev, lean, monkeypatch, warm, ckpt=ckpt)
        # (warp-major is compiled for the chunk-ring K1s: on the 1,203-basin grid the forced layout
        # sends the main forward to K1 — warm-up K1s + K2s remain)
        want = 0 if not lean else (2 if (layout == 1 and B == 1203) else 3)
        assert _lean_count() - n0 == want
        res[layout] = (out['streamflow'].clone(), grad.clone())
    if lean and B == 1203:      # different kernel families: fp32 contraction noise, not bit-identity
        assert_close(res[1][0], res[0][0], XTOL, 'layout 1 (K1 + K2s) vs layout 0 (K1s + K2s): streamflow')
        assert_close(res[1][1], res[0][1], XTOL, 'layout 1 (K1 + K2s) vs layout 0 (K1s + K2s): grad')
    else:
        assert torch.equal(res[0][0], res[1][0])
        assert torch.equal(res[0][1], res[1][1])
    assert torch.isfinite(res[1][1]).all() and float(res[1][1].abs().max()) > 0


def test_state_store_layout_policy():
    """Warp-major above the stage-pipelined regime for the lean-served sets; planes for the
    all-dynamic (TMA-staged) set, small grids and the hbv_2 state series."""
    from hydrodl2_b200 import ops
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 0, 'dynamic_params': {'Hbv': D2}, 'nmul': NMUL}, device=dev)
    spec = m._spec(D2, True)
    assert ops._ckpt_layout(spec, 2500, 1) == 1 and ops._ckpt_layout(spec, 22500, 4) == 1
    assert ops._ckpt_layout(spec, 531, 1) == 0 and ops._ckpt_layout(spec, 2000, 1) == 0 and ops._ckpt_layout(spec, 2500, 2) == 0
    M11 = hydrodl2.load_model('hbv_1_1p', ver_name='Hbv_1_1p')
    names = list(M11({'dynamic_params': {'Hbv_1_1p': []}, 'nmul': NMUL}, device=dev).parameter_bounds.keys())
    m11 = M11({'warm_up': 0, 'dynamic_params': {'Hbv_1_1p': names}, 'nmul': NMUL}, device=dev)
    assert ops._ckpt_layout(m11._spec(names, True), 22500, 1) == 0
    _cabi.set_option('ckpt_layout', 0)
    assert ops._ckpt_layout(spec, 2500, 1) == 0
