"""The clean-gradient-plane cache (ops._clean_plane): on small grids the dense parameter-gradient
plane is handed out again without its 488 MB memset when that is provably equivalent — nobody else
references its storage, it was not modified in place, same run plan.  Every case compares with the
cache switched off (a fresh zeroed plane every step): bit-identical gradients."""

import pytest
import torch

pytestmark = pytest.mark.gpu

NMUL = 16
D2 = ['parBETA', 'parBETAET']


def _setup(T=40, B=37, warm=6, seed=5):
    import hydrodl2_b200 as hydrodl2
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': D2}, 'nmul': NMUL}, device=dev)
    x = O.synthetic_forcing(T, B, seed=seed).to(dev)
    g = torch.Generator().manual_seed(seed + 1)
    ps = [torch.randn(T, B, 13 * NMUL + 2, generator=g).to(dev) for _ in range(3)]
    return m, x, ps


def _grad(m, x, p, key='streamflow'):
    pg = p.clone().requires_grad_(True)
    out = m({'x_phy': x}, pg)
    out[key].sum().backward()
    return pg.grad


def _reference_grads(m, x, ps, keys):
    from hydrodl2_b200 import ops
    ops.REUSE_GRAD_PLANE = False
    try:
        return [_grad(m, x, p, k).clone() for p, k in zip(ps, keys)]
    finally:
        ops.REUSE_GRAD_PLANE = True


def test_plane_is_reused_when_released_and_results_are_identical():
    from hydrodl2_b200 import ops
    m, x, ps = _setup()
    keys = ['streamflow'] * 3
    ref = _reference_grads(m, x, ps, keys)
    ops._PLANES.clear()
    ptrs = []
    for p, r in zip(ps, ref):
        g = _grad(m, x, p)
        assert torch.equal(g, r)
        ptrs.append(g.data_ptr())
        del g                                   # the gradient is released before the next step
    assert ptrs[0] == ptrs[1] == ptrs[2], 'the same plane should have been handed out again'
    assert sum(len(v) for v in ops._PLANES.values()) == 1


def test_plane_is_not_reused_while_the_gradient_or_a_view_of_it_is_alive():
    from hydrodl2_b200 import ops
    m, x, ps = _setup(seed=7)
    ref = _reference_grads(m, x, ps, ['streamflow'] * 3)
    ops._PLANES.clear()
    g0 = _grad(m, x, ps[0])
    g1 = _grad(m, x, ps[1])                     # g0 still alive: a different plane
    assert g1.data_ptr() != g0.data_ptr()
    assert torch.equal(g0, ref[0]) and torch.equal(g1, ref[1])
    view = g0[-1]                               # a view keeps the storage referenced
    del g0
    g2 = _grad(m, x, ps[2])
    assert g2.data_ptr() != view.data_ptr() - view.storage_offset() * 4
    assert torch.equal(g2, ref[2]) and torch.equal(view, ref[0][-1])


def test_plane_modified_in_place_is_not_reused():
    from hydrodl2_b200 import ops
    m, x, ps = _setup(seed=9)
    ref = _reference_grads(m, x, ps, ['streamflow'] * 3)
    ops._PLANES.clear()
    g0 = _grad(m, x, ps[0])
    ptr0 = g0.data_ptr()
    g0.add_(1.0)                                # e.g. weight decay added into .grad
    del g0
    g1 = _grad(m, x, ps[1])
    assert torch.equal(g1, ref[1])              # (a reused plane would carry the +1 in the zero columns)
    assert g1.data_ptr() != ptr0 or float(g1[0, 0, 20]) == 0.0


def test_routing_columns_are_zero_when_the_loss_skips_routing():
    """Step 1: loss on the routed flow writes the routing gradient into the last row; step 2 on
    the same (released) plane with a loss on an un-routed series must leave zeros there."""
    from hydrodl2_b200 import ops
    m, x, ps = _setup(seed=11)
    keys = ['streamflow', 'streamflow_no_rout', 'streamflow']
    ref = _reference_grads(m, x, ps, keys)
    ops._PLANES.clear()
    for p, k, r in zip(ps, keys, ref):
        g = _grad(m, x, p, k)
        assert torch.equal(g, r), k
        if k == 'streamflow_no_rout':
            assert float(g[-1, :, 13 * NMUL:].abs().max()) == 0.0
        del g


def test_plane_cache_under_cuda_graph():
    """The graph of a step replays into a pinned cached plane without a memset node; eager steps
    afterwards get planes of their own."""
    from hydrodl2_b200 import ops
    from hydrodl2_b200.graphs import GraphedStep
    m, x, ps = _setup(seed=13)
    ref = _reference_grads(m, x, ps[:1], ['streamflow'])
    ops._PLANES.clear()
    pg = ps[0].clone().requires_grad_(True)

    def step():
        pg.grad = None
        out = m({'x_phy': x}, pg)
        out['streamflow'].sum().backward()
        return pg.grad

    gs = GraphedStep(step, warmup=3, device=x.device)
    for _ in range(3):
        g = gs.replay()
    torch.cuda.synchronize()
    assert torch.equal(g, ref[0])
    pinned = [e for v in ops._PLANES.values() for e in v if e[3]]
    assert len(pinned) == 1
    g_eager = _grad(m, x, ps[0])
    assert torch.equal(g_eager, ref[0]) and g_eager.data_ptr() != g.data_ptr()


def test_fused_zero_fill_plane_is_kept_and_reused_on_large_grids():
    """Above the small-grid regime the adjoint zeroes its rows itself; the plane it initialised is
    kept, and the next step writes the gradient entries only (no zero fill, no memset) — same
    gradients bit for bit, with and without a warm-up period."""
    from hydrodl2_b200 import ops
    for warm in (5, 0):
        m, x, ps = _setup(T=23, B=2501, warm=warm, seed=17 + warm)
        ref = _reference_grads(m, x, ps, ['streamflow'] * 3)
        ops.release_grad_planes()
        ptrs = []
        for p, r in zip(ps, ref):
            g = _grad(m, x, p)
            assert torch.equal(g, r)
            ptrs.append(g.data_ptr())
            del g
        assert ptrs[0] == ptrs[1] == ptrs[2]
        g = _grad(m, x, ps[0], 'streamflow_no_rout')       # no routed cotangent: routing columns zeroed
        assert float(g[-1, :, 13 * NMUL:].abs().max()) == 0.0
        del g
    ops.release_grad_planes()
    assert all(e[3] for v in ops._PLANES.values() for e in v)      # (planes a CUDA graph replays into stay)


@pytest.mark.parametrize('warm', [0, 9])
def test_hbv_adj_plane_is_kept_and_reused(warm):
    """The implicit scheme's adjoint (K3) zero-fills its rows itself; the plane it initialised is
    kept and the next steps write gradient entries only — bit-identical gradients."""
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200 import ops
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B = 31, 21
    M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
    m = M({'warm_up': warm, 'dynamic_params': {'HbvAdj': D2}, 'nmul': NMUL}, device=dev)
    x = O.synthetic_forcing(T, B, seed=23).to(dev)
    g = torch.Generator().manual_seed(24)
    ps = [torch.randn(T, B, 13 * NMUL + 2, generator=g).to(dev) for _ in range(3)]

    def grad(p):
        pg = p.clone().requires_grad_(True)
        m({'x_phy': x}, pg)['flow_sim'].sum().backward()
        return pg.grad

    ops.REUSE_GRAD_PLANE = False
    try:
        ref = [grad(p).clone() for p in ps]
    finally:
        ops.REUSE_GRAD_PLANE = True
    ops.release_grad_planes()
    ptrs = []
    for p, r in zip(ps, ref):
        gg = grad(p)
        assert torch.equal(gg, r)
        ptrs.append(gg.data_ptr())
        del gg
    assert ptrs[0] == ptrs[1] == ptrs[2]
    ops.release_grad_planes()
