"""GPU parity: CUDA path (through the C-ABI) vs the reference's own outputs (golden
fixtures produced by tests/golden/make_golden.py from the unmodified reference) and vs the
CPU oracle on seeded inputs.  Tolerances: max-norm relative 1e-5 fluxes/states, 1e-4 grads."""

import pytest
import torch

from hydrodl2_b200 import _cabi
from conftest import RTOL_FLUX, RTOL_GRAD, STATE_FLOOR, assert_close, assert_grad_close, flux_rtol, load_golden

pytestmark = pytest.mark.gpu

PACKED = ['hbv_static', 'hbv_d2', 'hbv_d2_drop_nowarm', 'hbv_1_1p_d3', 'hbv_1_1p_d14']
CLS = {'hbv': 'Hbv', 'hbv_1_1p': 'Hbv_1_1p'}


def _run_packed(g, dev, ckpt=16):
    import hydrodl2_b200 as hydrodl2
    model = str(g['model'])
    cls = CLS[model]
    T, B, nmul, warm_up, seed = (int(v) for v in g['meta'])
    M = hydrodl2.load_model(model, ver_name=cls)
    cfg = {'warm_up': warm_up, 'dynamic_params': {cls: [str(s) for s in g['dyn']]}, 'nmul': nmul,
           'dy_drop': float(g['dy_drop']), 'warm_up_states': bool(int(g['warm_up_states'])),
           'ckpt_interval': ckpt}
    m = M(cfg, device=dev)
    x = g['x_phy'].to(dev)
    p = g['parameters'].to(dev).requires_grad_(True)
    torch.manual_seed(seed)
    out = m({'x_phy': x}, p)
    return m, out, p


@pytest.mark.parametrize('case', PACKED)
def test_forward_matches_reference(case):
    dev = torch.device('cuda:0')
    g = load_golden(case)
    m, out, _ = _run_packed(g, dev)
    assert set(out.keys()) == set(g['out'].keys())
    for k, ref in g['out'].items():
        assert_close(out[k], ref, RTOL_FLUX, f'{case}:{k}')
    for name, s in zip(m.state_names, m.get_states()):
        assert_close(s, g['states'][name], RTOL_FLUX, f'{case}:state {name}')


@pytest.mark.parametrize('ckpt', [1, 8, 16, 32])
@pytest.mark.parametrize('case', PACKED)
def test_gradient_matches_reference(case, ckpt):
    dev = torch.device('cuda:0')
    g = load_golden(case)
    m, out, p = _run_packed(g, dev, ckpt)
    loss = 0.0
    for k, c in g['cot'].items():
        loss = loss + (out[k] * c.to(dev)).sum()
    loss.backward()
    assert_grad_close(p.grad, g['grad_parameters'], f'{case}:grad K={ckpt}', int(g['meta'][2]))


@pytest.mark.parametrize('ring', ['0', '1'])
@pytest.mark.parametrize('ckpt', [1, 16])
@pytest.mark.parametrize('case', PACKED)
def test_gradient_both_input_paths(case, ckpt, ring, monkeypatch):
    """The golden cases are small, so by default they take the shared-memory-ring kernels; force
    each input path (HBV_B200_RING: 0 = register prefetch of the throughput regime, incl. its
    every-state-stored sweep at K = 1; 1 = cp.async ring) and check fluxes + gradient on both."""
    _cabi.set_option('ring', int(ring))
    dev = torch.device('cuda:0')
    g = load_golden(case)
    m, out, p = _run_packed(g, dev, ckpt)
    loss = 0.0
    for k, c in g['cot'].items():
        loss = loss + (out[k] * c.to(dev)).sum()
    loss.backward()
    for k, ref in g['out'].items():
        assert_close(out[k], ref, RTOL_FLUX, f'{case}:{k} ring={ring}')
    assert_grad_close(p.grad, g['grad_parameters'], f'{case}:grad K={ckpt} ring={ring}', int(g['meta'][2]))


def test_streamflow_only_gradient_vs_oracle():
    """Typical training use: loss on streamflow only (other flux grads are None)."""
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, B, nmul, warm = 200, 37, 16, 40
    dyn = ['parBETA', 'parBETAET']
    x = O.synthetic_forcing(T, B, seed=11)
    gen = torch.Generator().manual_seed(12)
    p = torch.randn(T, B, 13 * nmul + 2, generator=gen)
    pc = p.clone().requires_grad_(True)
    ref, S = O.forward_packed('hbv', x, pc, nmul=nmul, warm_up=warm, dynamic_params=dyn)
    ref['streamflow'].sum().backward()
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': dyn}, 'nmul': nmul}, device=dev)
    pg = p.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev)}, pg)
    out['streamflow'].sum().backward()
    for k in ref:
        assert_close(out[k], ref[k], RTOL_FLUX, k)
    assert_grad_close(pg.grad, pc.grad, 'grad', nmul)


@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('case', ['hbv_d2', 'hbv_d2_drop_nowarm', 'hbv_1_1p_d14'])
def test_fused_zero_fill_gradient(case, fused):
    """K2 writing the whole dense gradient tensor itself (gdyn_zero_fill = 1) into
    uninitialised memory gives the same gradient as the memset path."""
    from hydrodl2_b200 import ops
    dev = torch.device('cuda:0')
    g = load_golden(case)
    prev, ops.FUSED_ZERO_FILL = ops.FUSED_ZERO_FILL, fused   # default None = by dynamic share
    try:
        torch.empty(1 << 22, device=dev).fill_(float('nan'))   # poison the allocator's free blocks
        m, out, p = _run_packed(g, dev)
        loss = 0.0
        for k, c in g['cot'].items():
            loss = loss + (out[k] * c.to(dev)).sum()
        loss.backward()
    finally:
        ops.FUSED_ZERO_FILL = prev
    assert_grad_close(p.grad, g['grad_parameters'], f'{case}:grad fused zero-fill={fused}', int(g['meta'][2]))


# the last two: warm_up > 0 with warm_up_states=False on a split model, and the hourly
# model's per-unit 72-tap gamma-UH routing under the pair routing (hbv_2_hourly.py:684-705)
SPLIT = ['hbv_2_d3', 'hbv_2_d3_rout', 'hbv_2_hourly_d3', 'hbv_2_d3_nowarm', 'hbv_2_hourly_rout72']
SPLIT_CLS = {'hbv_2': 'Hbv_2', 'hbv_2_hourly': 'Hbv_2_hourly'}


def _run_split(g, dev, ckpt=16, state_series=True):
    import hydrodl2_b200 as hydrodl2
    model = str(g['model'])
    cls = SPLIT_CLS[model]
    T, B, nmul, _, seed = (int(v) for v in g['meta'])
    M = hydrodl2.load_model(model, ver_name=cls)
    cfg = {'dynamic_params': {cls: [str(s) for s in g['dyn']]}, 'nmul': nmul,
           'dy_drop': float(g['dy_drop']), 'routing': bool(int(g['routing'])),
           'ckpt_interval': ckpt, 'state_series': state_series}
    if 'warm_up' in g:
        cfg.update(warm_up=int(g['warm_up']), warm_up_states=bool(int(g['warm_up_states'])))
    m = M(cfg, device=dev)
    p0 = g['p0'].to(dev).requires_grad_(True)
    p1 = g['p1'].to(dev).requires_grad_(True)
    params = [p0, p1]
    xd = {'x_phy': g['x_phy'].to(dev), 'ac_all': g['ac_all'].to(dev), 'elev_all': g['elev_all'].to(dev)}
    if model == 'hbv_2_hourly':
        params.append(g['p2'].to(dev).requires_grad_(True))
        xd['outlet_topo'] = g['outlet_topo'].to(dev)
        xd['areas'] = g['areas'].to(dev)
    torch.manual_seed(seed)
    out = m(xd, params)
    return m, out, params


@pytest.mark.parametrize('ckpt', [16, 0])     # 0 = auto (K = 1): the state series aliases the stored states
@pytest.mark.parametrize('case', SPLIT)
def test_split_forward_matches_reference(case, ckpt):
    dev = torch.device('cuda:0')
    g = load_golden(case)
    m, out, _ = _run_split(g, dev, ckpt)
    assert set(out.keys()) == set(g['out'].keys())
    for k, ref in g['out'].items():
        assert_close(out[k], ref, flux_rtol(k), f'{case}:{k}')
    for name, s in zip(m.state_names, m._state_cache):
        assert_close(s, g['series'][name], RTOL_FLUX, f'{case}:series {name}', floor=STATE_FLOOR)


@pytest.mark.parametrize('ckpt,series', [(1, False), (16, True), (0, True), (1, True)])
@pytest.mark.parametrize('case', SPLIT)
def test_split_gradient_matches_reference(case, ckpt, series):
    dev = torch.device('cuda:0')
    g = load_golden(case)
    m, out, params = _run_split(g, dev, ckpt, state_series=series)
    loss = 0.0
    for k, c in g['cot'].items():
        loss = loss + (out[k] * c.to(dev)).sum()
    loss.backward()
    assert_grad_close(params[0].grad, g['grad']['p0'], f'{case}:grad dyn K={ckpt}', int(g['meta'][2]))
    assert_grad_close(params[1].grad, g['grad']['p1'], f'{case}:grad static K={ckpt}', int(g['meta'][2]))
    if len(params) > 2:
        assert_grad_close(params[2].grad, g['grad']['p2'], f'{case}:grad distr K={ckpt}', 1)
    if series:
        for name, s in zip(m.state_names, m._state_cache):
            assert_close(s, g['series'][name], RTOL_FLUX, f'{case}:series {name} K={ckpt}', floor=STATE_FLOOR)


def test_mts_matches_reference():
    """Hbv_2_mts training path: daily warm-up -> state hand-over -> hourly run (golden from the
    reference's own Hbv_2_mts.forward)."""
    import hydrodl2_b200 as hydrodl2
    from test_oracle_golden import _mts_inputs
    dev = torch.device('cuda:0')
    g = load_golden('hbv_2_mts_train')
    nmul = int(g['meta'][3])
    dyn = [str(s) for s in g['dyn']]
    M = hydrodl2.load_model('hbv_2_mts', ver_name='Hbv_2_mts')
    lo_cfg = {'dynamic_params': {'Hbv_2': dyn}, 'nmul': nmul, 'cache_states': True}
    hi_cfg = {'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul,
              'train_spatial_chunk_size': 10 ** 6, 'simulate_spatial_chunk_size': 10 ** 6,
              'simulate_temporal_chunk_size': 10 ** 6, 'train_warmup': 0}
    m = M(lo_cfg, hi_cfg, device=dev)
    xd, p, params = _mts_inputs(g, dev)
    out = m(xd, params)
    assert set(out.keys()) == {'Qs'}
    assert_close(out['Qs'], g['out']['Qs'], RTOL_FLUX, 'mts:Qs')
    (out['Qs'] * g['cot']['Qs'].to(dev)).sum().backward()
    assert p['lo_dyn'].grad is None or float(p['lo_dyn'].grad.abs().max()) == 0.0
    for k in ('lo_sta', 'hi_dyn', 'hi_sta'):
        assert_close(p[k].grad, g['grad'][k], RTOL_GRAD, f'mts:grad {k}')


def test_mts_chunked_simulation_equals_unchunked():
    """Spatial + temporal chunking (hbv_2_mts.py:204-279; broken in the reference) gives the same
    runoff as the un-chunked run and the same routed flow as one un-chunked routing call."""
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200.routing import distr_routing
    from test_oracle_golden import _mts_inputs
    dev = torch.device('cuda:0')
    g = load_golden('hbv_2_mts_train')
    nmul = int(g['meta'][3])
    dyn = [str(s) for s in g['dyn']]
    M = hydrodl2.load_model('hbv_2_mts', ver_name='Hbv_2_mts')
    lo_cfg = {'dynamic_params': {'Hbv_2': dyn}, 'nmul': nmul, 'cache_states': True}

    def run(spatial, temporal, simulate):
        hi_cfg = {'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul,
                  'train_spatial_chunk_size': spatial, 'simulate_spatial_chunk_size': spatial,
                  'simulate_temporal_chunk_size': temporal, 'train_warmup': 0}
        m = M(lo_cfg, hi_cfg, device=dev)
        m.set_mode(simulate)
        xd, _, params = _mts_inputs(g, dev)
        with torch.no_grad():
            return m(xd, params), xd, params

    full, xd, params = run(10 ** 6, 10 ** 6, False)
    chunked, _, _ = run(4, 10 ** 6, True)
    assert_close(chunked['Qs'], full['Qs'], 1e-6, 'chunked Qs')
    ref_rout = distr_routing(full['Qs'], params[1][2], xd['outlet_topo'], xd['areas'])
    assert_close(chunked['streamflow'], ref_rout, 1e-6, 'chunked streamflow')


def test_no_cpu_fallback():
    import hydrodl2_b200 as hydrodl2
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'dynamic_params': {'Hbv': []}, 'nmul': 2}, device=torch.device('cpu'))
    with pytest.raises(RuntimeError, match='no CPU path'):
        m({'x_phy': torch.zeros(4, 3, 3)}, torch.zeros(4, 3, 26))
