"""CPU: the oracle (oracle/hbv_oracle.py) is pinned against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  float32 must agree essentially bit for bit
(same torch ops in the same order); autograd gradients of the oracle must match the
reference's autograd gradients."""

import pytest
import torch

from conftest import assert_close, load_golden
from oracle import hbv_oracle as O

PACKED = ['hbv_static', 'hbv_d2', 'hbv_d2_drop_nowarm', 'hbv_1_1p_d3', 'hbv_1_1p_d14']
SPLIT = ['hbv_2_d3', 'hbv_2_d3_rout', 'hbv_2_hourly_d3', 'hbv_2_d3_nowarm', 'hbv_2_hourly_rout72']
TIGHT = 1e-7


def _packed(g, dtype=None):
    T, B, nmul, warm_up, seed = (int(v) for v in g['meta'])
    p = g['parameters'].clone().requires_grad_(True)
    torch.manual_seed(seed)
    out, S = O.forward_packed(str(g['model']), g['x_phy'], p, nmul=nmul, warm_up=warm_up,
                              dynamic_params=[str(s) for s in g['dyn']],
                              dy_drop=float(g['dy_drop']),
                              warm_up_states=bool(int(g['warm_up_states'])), dtype=dtype)
    return out, S, p


@pytest.mark.parametrize('case', PACKED)
def test_packed_forward_and_grad(case):
    g = load_golden(case)
    out, S, p = _packed(g)
    assert set(out) == set(g['out'])
    for k, ref in g['out'].items():
        assert_close(out[k], ref, TIGHT, f'{case}:{k}')
    for name, s in zip(['SNOWPACK', 'MELTWATER', 'SM', 'SUZ', 'SLZ'], S):
        assert_close(s, g['states'][name], TIGHT, f'{case}:{name}')
    loss = sum((out[k] * c).sum() for k, c in g['cot'].items())
    loss.backward()
    assert_close(p.grad, g['grad_parameters'], 1e-6, f'{case}:grad')


@pytest.mark.parametrize('case', SPLIT)
def test_split_forward_and_grad(case):
    g = load_golden(case)
    T, B, nmul, _, seed = (int(v) for v in g['meta'])
    model = str(g['model'])
    p0 = g['p0'].clone().requires_grad_(True)
    p1 = g['p1'].clone().requires_grad_(True)
    params = [p0, p1]
    xd = {'x_phy': g['x_phy'], 'ac_all': g['ac_all'], 'elev_all': g['elev_all']}
    if model == 'hbv_2_hourly':
        p2 = g['p2'].clone().requires_grad_(True)
        params.append(p2)
        xd['outlet_topo'] = g['outlet_topo']
        xd['areas'] = g['areas']
    torch.manual_seed(seed)
    out, series = O.forward_split(model, xd, params, nmul=nmul,
                                  dynamic_params=[str(s) for s in g['dyn']],
                                  dy_drop=float(g['dy_drop']), routing=bool(int(g['routing'])))
    for k, ref in g['out'].items():
        assert_close(out[k], ref, TIGHT, f'{case}:{k}')
    for name, s in zip(['SNOWPACK', 'MELTWATER', 'SM', 'SUZ', 'SLZ'], series):
        assert_close(s, g['series'][name], TIGHT, f'{case}:{name}')
    loss = sum((out[k] * c).sum() for k, c in g['cot'].items())
    loss.backward()
    assert_close(p0.grad, g['grad']['p0'], 1e-6, f'{case}:grad p0')
    assert_close(p1.grad, g['grad']['p1'], 1e-6, f'{case}:grad p1')
    if model == 'hbv_2_hourly':
        assert_close(params[2].grad, g['grad']['p2'], 1e-6, f'{case}:grad p2')


def _mts_inputs(g, dev='cpu'):
    xd = {'x_phy_low_freq': g['x_low'].to(dev), 'x_phy_high_freq': g['x_high'].to(dev),
          'ac_all': g['ac_all'].to(dev), 'elev_all': g['elev_all'].to(dev),
          'outlet_topo': g['outlet_topo'].to(dev), 'areas': g['areas'].to(dev)}
    p = {k: g[k].to(dev).clone().requires_grad_(True) for k in ('lo_dyn', 'lo_sta', 'hi_dyn', 'hi_sta')}
    params = ([p['lo_dyn'], p['lo_sta']], [p['hi_dyn'], p['hi_sta'], g['hi_distr'].to(dev)])
    return xd, p, params


def test_mts_forward_and_grad():
    g = load_golden('hbv_2_mts_train')
    nmul = int(g['meta'][3])
    dyn = [str(s) for s in g['dyn']]
    xd, p, params = _mts_inputs(g)
    out, _ = O.forward_mts(xd, params, nmul=nmul, low_dynamic=dyn, high_dynamic=dyn)
    assert_close(out['Qs'], g['out']['Qs'], TIGHT, 'mts:Qs')
    (out['Qs'] * g['cot']['Qs']).sum().backward()
    assert p['lo_dyn'].grad is None                      # warm-up states are detached
    for k in ('lo_sta', 'hi_dyn', 'hi_sta'):
        assert_close(p[k].grad, g['grad'][k], 1e-6, f'mts:grad {k}')


def test_float64_arbiter_is_close_to_float32():
    g = load_golden('hbv_d2')
    out64, _, _ = _packed(g, dtype=torch.float64)
    for k, ref in g['out'].items():
        assert_close(out64[k].float(), ref, 1e-4, f'fp64:{k}')


def test_uh_conv_definition():
    """uh_conv == explicit causal sum y[t] = sum_k UH[k] x[t-k] (uh_routing.py:25-57)."""
    torch.manual_seed(0)
    x = torch.rand(7, 1, 40)
    a = torch.rand(40, 7, 1) * 2.9
    b = torch.rand(40, 7, 1) * 6.5
    UH = O.uh_gamma(a, b, lenF=15).permute(1, 2, 0)
    assert UH.shape == (7, 1, 15)
    assert torch.allclose(UH.sum(-1), torch.ones(7, 1), atol=1e-6)
    assert torch.allclose(O.uh_conv(x, UH), O.uh_conv_direct(x, UH), atol=1e-6)


def test_uh_short_series_truncates_lenF():
    a = torch.rand(5, 3, 1)
    b = torch.rand(5, 3, 1)
    assert O.uh_gamma(a, b, lenF=15).shape[0] == 5
