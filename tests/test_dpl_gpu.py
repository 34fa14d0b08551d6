"""a parameter network in front of the CUDA path trains end to end — gradients reach
the network through the hand-written adjoint, and a few optimiser steps reduce the loss."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dpl_training_reduces_loss():
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200.dpl import DplModel, allreduce_gradients, rmse_loss
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, W = 120, 24, 30
    x_phy = O.synthetic_forcing(T, B, seed=5).to(dev)
    attrs = torch.randn(1, B, 6, generator=torch.Generator().manual_seed(6)).expand(T, B, 6).to(dev)
    x_nn = torch.cat([(x_phy - x_phy.mean((0, 1))) / x_phy.std((0, 1)), attrs], dim=-1).contiguous()
    xd = {'x_phy': x_phy, 'xc_nn_norm': x_nn}
    cfg = {'warm_up': W, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': 16}
    Hbv = hydrodl2.load_model('hbv', ver_name='Hbv')
    torch.manual_seed(1)
    truth = DplModel(Hbv(cfg, device=dev), nx=9, hidden_size=32, dropout=0.0).to(dev)
    with torch.no_grad():
        obs = truth(xd)['streamflow'].clone()
    obs[::7] = float('nan')                                   # missing observations are masked
    torch.manual_seed(2)
    model = DplModel(Hbv(cfg, device=dev), nx=9, hidden_size=32, dropout=0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    losses = []
    for _ in range(25):
        opt.zero_grad(set_to_none=True)
        loss = rmse_loss(model(xd)['streamflow'], obs)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        assert allreduce_gradients(model.parameters()) == sum(p.numel() for p in model.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0], losses
