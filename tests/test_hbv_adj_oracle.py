"""CPU checks of the implicit-scheme oracle (oracle/hbv_adj_oracle.py).

hbv_adj has no runnable reference (parity unpinned), so the oracle is pinned
by its own mathematics instead: the Newton solution satisfies the backward-Euler residual, and
the adjoint gradient equals a float64 finite difference of the forward solve."""

import torch

from oracle import hbv_adj_oracle as A
from oracle import hbv_oracle as O


def _inputs(T, B, nmul, seed, n_par=13):
    x = O.synthetic_forcing(T, B, seed=seed).double()
    p = torch.randn(T, B, n_par * nmul + 2, generator=torch.Generator().manual_seed(seed + 1)).double()
    return x, p


def test_newton_solution_satisfies_residual():
    T, B, nmul = 40, 5, 4
    dyn = ['parBETA', 'parBETAET']
    x, p = _inputs(T, B, nmul, 5)
    bounds = A.adj_bounds(dyn)
    out = A.forward_adj(x, p, nmul=nmul, dynamic_params=dyn, tol=1e-9, max_updates=30, return_states=True)
    ys = out['states'].permute(1, 3, 2, 0).reshape(T, B * nmul, 5)          # [T, N(j*B+b), 5]
    phy = torch.sigmoid(p[:, :, :13 * nmul]).view(T, B, 13, nmul).permute(0, 3, 1, 2).reshape(T, B * nmul, 13)
    th = phy[-1].unsqueeze(0).repeat(T, 1, 1).clone()
    for i in (0, 12):
        th[:, :, i] = phy[:, :, i]
    clim = x.unsqueeze(1).repeat(1, nmul, 1, 1).view(T, B * nmul, 3)
    xt = torch.zeros(B * nmul, 5, dtype=torch.float64)
    for t in range(T):
        g = A.residual(ys[t], th[t], xt, clim[t], 1.0, bounds)
        assert g.abs().max().item() < 1e-8, (t, g.abs().max().item())
        xt = ys[t]


def test_adjoint_gradient_matches_finite_difference():
    T, B, nmul, warm = 30, 3, 2, 8
    dyn = ['parBETA', 'parBETAET']
    x, p = _inputs(T, B, nmul, 9)
    kw = dict(nmul=nmul, warm_up=warm, dynamic_params=dyn, tol=1e-10, max_updates=40)
    pp = p.clone().requires_grad_(True)
    w = torch.randn(T - warm, B, 1, generator=torch.Generator().manual_seed(3)).double()
    loss = (A.forward_adj(x, pp, **kw)['flow_sim'] * w).sum()
    loss.backward()
    g = pp.grad
    # directional finite differences along random directions (kinks make single coordinates fragile)
    gen = torch.Generator().manual_seed(4)
    for _ in range(3):
        d = torch.randn(p.shape, generator=gen).double()
        eps = 1e-6
        lp = (A.forward_adj(x, p + eps * d, **kw)['flow_sim'] * w).sum()
        lm = (A.forward_adj(x, p - eps * d, **kw)['flow_sim'] * w).sum()
        fd = ((lp - lm) / (2 * eps)).item()
        an = (g * d).sum().item()
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 1e-8, (fd, an)


def test_reference_schedule_is_within_its_own_tolerance_of_converged_solution():
    """The reference's schedule (global gtol 1e-3, <= 4 updates, lazy Jacobian) and the tightly
    converged per-lane solve agree to the looseness the reference accepts."""
    T, B, nmul = 60, 4, 4
    dyn = ['parBETA', 'parBETAET']
    x, p = _inputs(T, B, nmul, 13)
    a = A.forward_adj(x, p, nmul=nmul, dynamic_params=dyn, newton='reference')['flow_sim']
    b = A.forward_adj(x, p, nmul=nmul, dynamic_params=dyn, newton='lane')['flow_sim']
    assert (a - b).abs().max().item() <= 2e-2 * b.abs().max().item()


def test_twelve_parameter_form_runs():
    T, B, nmul = 20, 3, 2
    x, p = _inputs(T, B, nmul, 17, n_par=12)
    out = A.forward_adj(x, p, nmul=nmul, dynamic_params=['parBETA'])
    assert out['flow_sim'].shape == (T, B, 1) and torch.isfinite(out['flow_sim']).all()
