"""GPU parity of K3 (implicit HBV, csrc/hbv_adj.cu) through the HbvAdj drop-in class against
the float64 oracle restatement of hbv_adj.py (oracle/hbv_adj_oracle.py, same per-lane Newton
rule).  hbv_adj has no runnable reference: parity here is against the restatement (unpinned,
DESIGN.md §3).  Tolerances: max-norm relative 1e-5 flow, 1e-4 parameter gradient."""

import pytest
import torch

from conftest import RTOL_FLUX, RTOL_GRAD, assert_close

pytestmark = pytest.mark.gpu

CASES = [
    # name, T, B, nmul, warm_up, dynamic, n_par
    ('static', 120, 9, 4, 0, [], 12),
    ('d2_warm', 150, 7, 16, 40, ['parBETA', 'parBETAET'], 13),
    ('d2_nowarm', 100, 33, 8, 0, ['parBETA', 'parBETAET'], 13),
    ('all_dynamic', 90, 5, 4, 20, ['parBETA', 'parFC', 'parK0', 'parK1', 'parK2', 'parLP', 'parPERC',
                                   'parUZL', 'parTT', 'parCFMAX', 'parCFR', 'parCWH', 'parBETAET'], 13),
    ('d1_twelve', 80, 6, 2, 10, ['parK1'], 12),
]


def _run(case, routing=True):
    from oracle import hbv_adj_oracle as AO
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    name, T, B, nmul, warm, dyn, n_par = case
    dev = torch.device('cuda:0')
    x = O.synthetic_forcing(T, B, seed=21)
    gen = torch.Generator().manual_seed(22)
    p = torch.randn(T, B, n_par * nmul + 2, generator=gen)
    cot = torch.randn(T - warm, B, 1, generator=gen)
    p64 = p.double().requires_grad_(True)
    ref = AO.forward_adj(x.double(), p64, nmul=nmul, warm_up=warm, dynamic_params=dyn, routing=routing)
    (ref['flow_sim'] * cot.double()).sum().backward()
    M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
    m = M({'warm_up': warm, 'dynamic_params': {'HbvAdj': dyn}, 'nmul': nmul, 'routing': routing}, device=dev)
    pg = p.to(dev).requires_grad_(True)
    out = m({'x_phy': x.to(dev)}, pg)
    (out['flow_sim'] * cot.to(dev)).sum().backward()
    return m, out, pg, ref, p64


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_adj_flow_and_gradient(case):
    m, out, pg, ref, p64 = _run(case)
    assert_close(out['flow_sim'], ref['flow_sim'], RTOL_FLUX, f'{case[0]}:flow_sim')
    assert_close(pg.grad, p64.grad, RTOL_GRAD, f'{case[0]}:grad')
    stats = m.newton_stats.cpu().tolist()
    assert 1 <= stats[0] <= m.newton_max_updates and stats[1] == 0, stats


@pytest.mark.parametrize('B', [7, 64])
def test_adj_ring_adjoint_matches_register_form(B):
    """The standard case (nmul 16, [parBETA, parBETAET]) takes the cp.async-ring adjoint
    (hbv_adj_bwd_ring_kernel); option ring = 0 keeps the register-prefetch kernel.  Same step
    arithmetic (one device function), so the gradients agree to fp32 contraction noise; the oracle
    leg of the ring form is `test_adj_flow_and_gradient[d2_warm]`."""
    from hydrodl2_b200 import _cabi
    from oracle import hbv_oracle as O
    import hydrodl2_b200 as hydrodl2
    dev = torch.device('cuda:0')
    T, nmul, warm = 61, 16, 11
    x = O.synthetic_forcing(T, B, seed=31).to(dev)
    gen = torch.Generator().manual_seed(32)
    p = torch.randn(T, B, 13 * nmul + 2, generator=gen).to(dev)
    cot = torch.randn(T - warm, B, 1, generator=gen).to(dev)
    grads = {}
    for ring in (-1, 0):
        _cabi.set_option('ring', ring)
        M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
        m = M({'warm_up': warm, 'dynamic_params': {'HbvAdj': ['parBETA', 'parBETAET']}, 'nmul': nmul}, device=dev)
        pg = p.clone().requires_grad_(True)
        out = m({'x_phy': x}, pg)
        (out['flow_sim'] * cot).sum().backward()
        grads[ring] = pg.grad
    assert torch.isfinite(grads[-1]).all()
    assert_close(grads[-1], grads[0], 5e-6, f'K3 adjoint ring vs register form, B={B}')
    # structural zeros of the dense plane are exact zeros in both (fused zero fill)
    assert torch.equal(grads[-1] == 0, grads[0] == 0)


def test_adj_no_routing():
    case = CASES[1]
    m, out, pg, ref, p64 = _run(case, routing=False)
    assert_close(out['flow_sim'], ref['flow_sim'], RTOL_FLUX, 'flow_sim (no routing)')
    assert_close(pg.grad, p64.grad, RTOL_GRAD, 'grad (no routing)')


def test_adj_vs_reference_newton_schedule():
    """The (not runnable) reference stops its Newton loop on the BATCH-max residual, after at most
    4 updates, at the loose gtol = 1e-3 (hbv_adj.py:518-519,544); K3 iterates every lane to its own
    tolerance with up to 8 updates.  The oracle restates both schedules (newton='reference' /
    'lane'); this test states the gap between them and holds K3 to it: K3 must sit on the 'lane'
    side (1e-5) and no further from the reference schedule than that schedule is from the
    converged solution.  Measured on B200 (case d2_warm): the two schedules differ by 1.2e-4 (flow) and
    5.3e-3 (gradient) of max-norm; K3 is 1.7e-7 from the converged solution, the reference schedule 1.2e-4."""
    from oracle import hbv_adj_oracle as AO
    from oracle import hbv_oracle as O
    case = CASES[1]
    name, T, B, nmul, warm, dyn, n_par = case
    m, out, pg, ref_lane, p64 = _run(case)
    x = O.synthetic_forcing(T, B, seed=21)
    gen = torch.Generator().manual_seed(22)
    p = torch.randn(T, B, n_par * nmul + 2, generator=gen)
    cot = torch.randn(T - warm, B, 1, generator=gen)
    q64 = p.double().requires_grad_(True)
    ref_ref = AO.forward_adj(x.double(), q64, nmul=nmul, warm_up=warm, dynamic_params=dyn, newton='reference')
    (ref_ref['flow_sim'] * cot.double()).sum().backward()
    tight = AO.forward_adj(x.double(), p.double(), nmul=nmul, warm_up=warm, dynamic_params=dyn, tol=1e-10, max_updates=50)

    def err(a, b):
        return float((a.detach().double().cpu() - b.detach().double()).abs().max() / b.detach().double().abs().max())

    gap_flow = err(ref_ref['flow_sim'], ref_lane['flow_sim'])
    gap_grad = err(q64.grad, p64.grad)
    k3_flow = err(out['flow_sim'], ref_ref['flow_sim'])
    k3_grad = err(pg.grad, q64.grad)
    print(f'hbv_adj Newton schedules: reference-vs-lane oracle gap flow {gap_flow:.2e} grad {gap_grad:.2e}; '
          f'K3 vs reference schedule flow {k3_flow:.2e} grad {k3_grad:.2e}; '
          f'reference schedule vs converged {err(ref_ref["flow_sim"], tight["flow_sim"]):.2e}, '
          f'K3 vs converged {err(out["flow_sim"], tight["flow_sim"]):.2e}')
    assert k3_flow <= gap_flow + 1e-5 and k3_grad <= gap_grad + 1e-4
    # K3 (per-lane stop at the same tol, more updates allowed) is at least as close to the converged
    # solution as the reference's schedule
    assert err(out['flow_sim'], tight['flow_sim']) <= err(ref_ref['flow_sim'], tight['flow_sim']) + 1e-5


def test_adj_attributes_and_errors():
    import hydrodl2_b200 as hydrodl2
    M = hydrodl2.load_model('hbv_adj', ver_name='HbvAdj')
    m = M({'dynamic_params': {'HbvAdj': ['parBETAET']}, 'nmul': 16}, device=torch.device('cuda:0'))
    assert m.learnable_param_count == 13 * 16 + 2
    assert list(m.routing_parameter_bounds) == ['rout_a', 'rout_b']
    with pytest.raises(KeyError):
        M({'nmul': 4}, device=torch.device('cuda:0'))
    with pytest.raises(RuntimeError):
        m({'x_phy': torch.zeros(4, 2, 3)}, torch.zeros(4, 2, 210))   # CPU tensors: no CPU path
