"""A minimal differentiable-parameter-learning (dPL) wrapper around the HBV path.

δMG — the framework hydrodl2 plugs into — is not part of the reference repository; this module
provides just enough of it to run and time an END-TO-END training step without δMG:

    attributes + forcings -> parameter network -> raw parameters [T, B, ny]
                          -> Hbv.forward (this package: K1/K2/K4 kernels) -> streamflow
                          -> loss -> adjoint -> parameter-network backward -> gradient all-reduce

The parameter network is the LSTM head the reference's test configuration names
(`tests/config.yaml:51-55`: LSTM, hidden 256, dropout 0.5): linear-in, ReLU, one LSTM layer,
linear-out producing `learnable_param_count` raw values per (time step, basin) — the layout
`Hbv._unpack_parameters` expects (`hbv.py:182-215`).  It is ordinary `torch.nn` (cuDNN / cuBLAS):
plumbing around the hot path, not part of it.  Across GPUs the basins shard and the only
collective is ONE all-reduce of the flattened network gradients (`allreduce_gradients`).
"""

from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


class LstmParameterNetwork(torch.nn.Module):
    """[T, B, nx] normalised inputs -> [T, B, ny] raw (pre-sigmoid) HBV parameters."""

    def __init__(self, nx: int, ny: int, hidden_size: int = 256, dropout: float = 0.5) -> None:
        super().__init__()
        self.linear_in = torch.nn.Linear(nx, hidden_size)
        self.lstm = torch.nn.LSTM(hidden_size, hidden_size, num_layers=1)
        self.drop = torch.nn.Dropout(dropout)
        self.linear_out = torch.nn.Linear(hidden_size, ny)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = torch.relu(self.linear_in(x))
        h, _ = self.lstm(h)
        return self.linear_out(self.drop(h))


class DplModel(torch.nn.Module):
    """Parameter network + physical model, the pairing δMG's `DplModel` makes."""

    def __init__(self, phy_model: torch.nn.Module, nx: int, hidden_size: int = 256, dropout: float = 0.5) -> None:
        super().__init__()
        self.phy_model = phy_model
        self.nn_model = LstmParameterNetwork(nx, phy_model.learnable_param_count, hidden_size, dropout)

    def forward(self, x_dict: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """x_dict: 'x_phy' [T, B, 3] forcings for the physical model, 'xc_nn_norm' [T, B, nx]
        normalised forcings + static attributes for the network (δMG's key names)."""
        parameters = self.nn_model(x_dict['xc_nn_norm'])
        return self.phy_model(x_dict, parameters)


def rmse_loss(pred: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """Root-mean-square error over the non-missing observations."""
    mask = ~torch.isnan(obs)
    diff = torch.where(mask, pred - torch.nan_to_num(obs), torch.zeros_like(pred))
    return torch.sqrt((diff * diff).sum() / mask.sum().clamp(min=1))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True) -> int:
    """Sum (or average) the gradients of `params` over all ranks with ONE collective on a flat
    buffer; returns the number of elements reduced.  No-op for a single process."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    n = sum(g.numel() for g in grads)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return n
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return n
