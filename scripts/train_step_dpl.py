#!/usr/bin/env python
"""End-to-end dPL training step: LSTM parameter network -> Hbv (CUDA path) ->
RMSE loss -> adjoint -> network backward -> one all-reduce of the network gradients -> Adadelta.

    python scripts/train_step_dpl.py [--basins 531] [--rho 730] [--warm-up 365] [--steps 20]
    python -m torch.distributed.run --nproc-per-node N ... scripts/train_step_dpl.py

Prints one JSON object (rank 0): ms per training step, the share spent in the HBV kernels (CUDA
events around every C-ABI call) and the loss of the first / last step.  Synthetic forcings
 and attributes; "observations" are the streamflow of a hidden random network, so the
loss has something to fit.  Each rank trains on its own basins (weak scaling).
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--basins', type=int, default=531)
    ap.add_argument('--rho', type=int, default=730)
    ap.add_argument('--warm-up', type=int, default=365)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--n-attr', type=int, default=35)
    args = ap.parse_args()

    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200 import dist as D, ops
    from hydrodl2_b200.dpl import DplModel, allreduce_gradients, rmse_loss
    rank, local, world = D.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    T, B, W, nmul = args.warm_up + args.rho, args.basins, args.warm_up, 16
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    d = torch.arange(T, dtype=torch.float32, device=dev).view(T, 1)
    season = torch.sin(2 * math.pi * (d - 110) / 365)
    ob = torch.rand(1, B, generator=g, device=dev) * 16 - 8
    tmean = 5 + 12 * season + ob + 4 * torch.randn(T, B, generator=g, device=dev)
    prcp = 5 * torch.relu(torch.randn(T, B, generator=g, device=dev))
    pet = torch.relu(2 + 2 * season) + 0.5 * torch.rand(T, B, generator=g, device=dev)
    x_phy = torch.stack([prcp, tmean, pet], dim=-1).contiguous()
    attrs = torch.randn(1, B, args.n_attr, generator=g, device=dev).expand(T, B, args.n_attr)
    x_nn = torch.cat([(x_phy - x_phy.mean((0, 1))) / x_phy.std((0, 1)), attrs], dim=-1).contiguous()
    x_dict = {'x_phy': x_phy, 'xc_nn_norm': x_nn}

    Hbv = hydrodl2.load_model('hbv', ver_name='Hbv')
    cfg = {'warm_up': W, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': nmul}
    torch.manual_seed(7)                     # same initial weights on every rank
    model = DplModel(Hbv(cfg, device=dev), nx=x_nn.shape[-1]).to(dev)
    with torch.no_grad():                    # hidden "truth": another random network
        torch.manual_seed(8 + rank)
        truth = DplModel(Hbv(cfg, device=dev), nx=x_nn.shape[-1]).to(dev).eval()
        obs = truth(x_dict)['streamflow'].clone()
    torch.manual_seed(9 + rank)
    opt = torch.optim.Adadelta(model.parameters(), lr=1.0)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = rmse_loss(model(x_dict)['streamflow'], obs)
        loss.backward()
        allreduce_gradients(model.parameters())
        opt.step()
        return loss

    losses = [float(step()) for _ in range(3)]
    torch.cuda.synchronize(dev)
    D.barrier()
    ops.PROFILE = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        last = step()
    e1.record()
    torch.cuda.synchronize(dev)
    prof, ops.PROFILE = ops.PROFILE, None
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    hbv_ms = sum(sum(a.elapsed_time(b) for a, b in v) for v in prof.values()) / args.steps
    n_par = sum(p.numel() for p in model.parameters())
    if rank == 0:
        print(json.dumps({
            'what': f'dPL training step: LSTM(256) parameter network -> hbv (D2, nmul 16) -> RMSE -> adjoint -> '
                    f'network backward -> all-reduce of {n_par} gradients -> Adadelta',
            'basins_per_gpu': B, 'time_steps': T, 'warm_up': W, 'n_gpus': world,
            'ms_per_step': ms, 'hbv_kernels_ms_per_step': hbv_ms, 'hbv_share': hbv_ms / ms,
            'basin_timesteps_per_s': world * B * args.rho / (ms * 1e-3),
            'loss_first': losses[0], 'loss_last': float(last)}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
