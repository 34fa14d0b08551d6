"""SURVEY §8 f2 — stateful streaming: `cache_states` stepping and the hourly model's runoff
history buffer (hbv.py:321-324,358-359; hbv_2_hourly.py:766-796).

  * a run cut into chunks, each started from the storages the previous chunk cached, reproduces
    the un-routed fluxes of the one-shot run (the reference's own behaviour, SURVEY §5 [probe]);
    routed series are NOT expected to match — the daily UH convolution has no carry-over;
  * the hourly model fed one step at a time keeps <= 100 steps of runoff history in `_qs_buffer`
    and returns the last routed value: it must equal the oracle's pair routing of that window.
"""

import pytest
import torch

from conftest import RTOL_FLUX, STATE_FLOOR, assert_close

pytestmark = pytest.mark.gpu

UNROUTED = ['streamflow_no_rout', 'srflow_no_rout', 'ssflow_no_rout', 'gwflow_no_rout', 'AET_hydro', 'SWE',
            'recharge', 'excs', 'evapfactor', 'tosoil', 'percolation']


@pytest.mark.parametrize('model,cls,npar', [('hbv', 'Hbv', 13), ('hbv_1_1p', 'Hbv_1_1p', 14)])
def test_cache_states_chunked_equals_one_shot(model, cls, npar):
    import hydrodl2_b200 as hydrodl2
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul, chunk = 90, 21, 16, 30
    dyn = ['parBETA', 'parBETAET']
    x = O.synthetic_forcing(T, B, seed=71).to(dev)
    g = torch.Generator().manual_seed(72)
    p = torch.randn(1, B, npar * nmul + 2, generator=g).repeat(T, 1, 1)     # static values: the same at every row
    for i in (0, 12):                                                        # the two dynamic blocks vary in time
        p[:, :, i * nmul:(i + 1) * nmul] = torch.randn(T, B, nmul, generator=g)
    p = p.to(dev)
    M = hydrodl2.load_model(model, ver_name=cls)
    cfg = {'warm_up': 0, 'dynamic_params': {cls: dyn}, 'nmul': nmul}
    with torch.no_grad():
        one = M(cfg, device=dev)({'x_phy': x}, p)
        m = M(dict(cfg, cache_states=True), device=dev)
        parts = [m({'x_phy': x[t0:t0 + chunk].contiguous()}, p[t0:t0 + chunk].contiguous()) for t0 in range(0, T, chunk)]
    for k in UNROUTED + (['capillary'] if model == 'hbv_1_1p' else []):
        got = torch.cat([q[k] for q in parts], dim=0)
        assert_close(got, one[k], 1e-6, f'{model} chunked cache_states: {k}')
    # the cached storages are those of the one-shot run's end
    full = M(cfg, device=dev)
    with torch.no_grad():
        full({'x_phy': x}, p)
    for name, a, b in zip(m.state_names, m.get_states(), full.get_states()):
        assert_close(a, b, 1e-6, f'{model} cached state {name}', floor=STATE_FLOOR)
    # load_states(get_states()) round trip: the reference hands out a list, load_states wants a tuple
    m2 = M(dict(cfg, cache_states=True), device=dev)
    m2.load_states(tuple(m.get_states()))
    assert all(torch.equal(a, b) for a, b in zip(m2.states, m.get_states()))


def test_hourly_qs_buffer_streaming_matches_oracle_routing():
    import hydrodl2_b200 as hydrodl2
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul = 110, 6, 4
    dyn = ['parBETA', 'parK0', 'parBETAET']
    g = torch.Generator().manual_seed(81)
    x = O.synthetic_forcing(T, B, seed=82, hourly=True)
    p0 = torch.rand(T, B, 3 * nmul, generator=g)
    p1 = torch.rand(B, 16 * nmul, generator=g)
    topo = torch.zeros(2, B)
    topo[0, :4] = 1
    topo[1, 2:] = 1                       # units 2, 3 drain to both gages
    areas = torch.rand(B, generator=g) * 99 + 1
    p2 = torch.rand(int(topo.sum()), 3, generator=g)
    base = {'ac_all': (torch.rand(B, generator=g) * 5000).to(dev), 'elev_all': (torch.rand(B, generator=g) * 3500).to(dev),
            'outlet_topo': topo.to(dev), 'areas': areas.to(dev)}
    M = hydrodl2.load_model('hbv_2_hourly', ver_name='Hbv_2_hourly')
    cfg = {'dynamic_params': {'Hbv_2_hourly': dyn}, 'nmul': nmul, 'routing': False}
    params = [p0.to(dev), p1.to(dev), p2.to(dev)]
    with torch.no_grad():
        one = M(cfg, device=dev)(dict(base, x_phy=x.to(dev)), params)
        m = M(dict(cfg, cache_states=True), device=dev)
        qs_steps, flow_steps = [], []
        for t in range(T):
            o = m(dict(base, x_phy=x[t:t + 1].to(dev)), [params[0][t:t + 1].contiguous(), params[1], params[2]])
            qs_steps.append(o['Qs'])
            flow_steps.append(o['streamflow'])
            assert o['streamflow'].shape == (1, 2, 1)
            assert len(m._qs_buffer) == min(t + 1, m._max_history)
    qs = torch.cat(qs_steps, dim=0)
    assert_close(qs, one['Qs'], 1e-6, 'hourly stepping: Qs')
    v = O.variant('hbv_2_hourly', dyn)
    distr = {k: O.change_param_range(p2[:, i], bd) for i, (k, bd) in enumerate(v.distr_bounds.items())}
    qs_cpu = one['Qs'].cpu()
    for t in (0, 1, 50, 99, 100, T - 1):
        lo = max(0, t + 1 - m._max_history)
        ref = O.distr_routing(qs_cpu[lo:t + 1], distr, topo, areas, lenF=v.lenF)[-1:]
        assert_close(flow_steps[t], ref, RTOL_FLUX, f'hourly stepping: routed flow at step {t} (history {t + 1 - lo})')


ROUTED = ['streamflow', 'srflow', 'ssflow', 'gwflow']


@pytest.mark.parametrize('chunk', [30, 7, 1])
def test_uh_carry_over_makes_chunked_routing_equal_one_shot(chunk):
    """Extension of SURVEY f2 (`uh_carry_over`, off by default): with the last lenF - 1 steps of
    un-routed flow carried from call to call, the ROUTED series of a run stepped chunk by chunk
    equal the one-shot run — chunks longer than, shorter than, and much shorter than the 15-tap
    unit hydrograph.  Without the option they do not (the reference's behaviour)."""
    import hydrodl2_b200 as hydrodl2
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul = 60, 13, 16
    dyn = ['parBETA', 'parBETAET']
    x = O.synthetic_forcing(T, B, seed=81).to(dev)
    g = torch.Generator().manual_seed(82)
    p = torch.randn(1, B, 13 * nmul + 2, generator=g).repeat(T, 1, 1)
    for i in (0, 12):
        p[:, :, i * nmul:(i + 1) * nmul] = torch.randn(T, B, nmul, generator=g)
    p = p.to(dev)
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    cfg = {'warm_up': 0, 'dynamic_params': {'Hbv': dyn}, 'nmul': nmul}
    with torch.no_grad():
        one = M(cfg, device=dev)({'x_phy': x}, p)
        m = M(dict(cfg, cache_states=True, uh_carry_over=True), device=dev)
        parts = [m({'x_phy': x[t0:t0 + chunk].contiguous()}, p[t0:t0 + chunk].contiguous()) for t0 in range(0, T, chunk)]
        plain = M(dict(cfg, cache_states=True), device=dev)
        parts0 = [plain({'x_phy': x[t0:t0 + chunk].contiguous()}, p[t0:t0 + chunk].contiguous()) for t0 in range(0, T, chunk)]
    for k in ROUTED + UNROUTED:
        got = torch.cat([q[k] for q in parts], dim=0)
        assert_close(got, one[k], RTOL_FLUX, f'uh_carry_over, chunk {chunk}: {k}', floor=STATE_FLOOR)
    if chunk < T:
        got0 = torch.cat([q['streamflow'] for q in parts0], dim=0)
        assert float((got0 - one['streamflow']).abs().max()) > 1e-3 * float(one['streamflow'].abs().max())
    # a new starting point clears the history; training is refused
    m.load_states(tuple(m.get_states()))
    assert m._uh_hist is None
    with pytest.raises(RuntimeError, match='streaming-inference'):
        m({'x_phy': x[:5].contiguous()}, p[:5].contiguous().requires_grad_(True))


def test_uh_carry_over_hbv_2():
    """The same extension for the split daily model (`hbv_2`: routing parameters in the static
    tensor)."""
    import hydrodl2_b200 as hydrodl2
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul, chunk = 48, 9, 16, 5
    dyn = ['parBETA', 'parK0', 'parBETAET']
    g = torch.Generator().manual_seed(91)
    x = O.synthetic_forcing(T, B, seed=92).to(dev)
    p0 = torch.rand(T, B, 3 * nmul, generator=g).to(dev)
    p1 = torch.rand(B, 13 * nmul + 2, generator=g).to(dev)
    xd = {'ac_all': (torch.rand(B, generator=g) * 5000).to(dev), 'elev_all': (torch.rand(B, generator=g) * 3500).to(dev)}
    M = hydrodl2.load_model('hbv_2', ver_name='Hbv_2')
    cfg = {'warm_up': 0, 'dynamic_params': {'Hbv_2': dyn}, 'nmul': nmul, 'routing': True}
    with torch.no_grad():
        one = M(cfg, device=dev)(dict(xd, x_phy=x), [p0, p1])
        m = M(dict(cfg, cache_states=True, uh_carry_over=True), device=dev)
        parts = [m(dict(xd, x_phy=x[t0:t0 + chunk].contiguous()), [p0[t0:t0 + chunk].contiguous(), p1])
                 for t0 in range(0, T, chunk)]
    for k in ROUTED:
        got = torch.cat([q[k] for q in parts], dim=0)
        assert_close(got, one[k], RTOL_FLUX, f'hbv_2 uh_carry_over: {k}', floor=STATE_FLOOR)
