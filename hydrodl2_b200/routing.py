"""Distributed (gage, unit) pair routing — host side of K4' (csrc/pair_route.cu).

Drop-in for ``Hbv_2_hourly.distr_routing`` (hbv_2_hourly.py:800-855): same arguments and result
``[T, n_gages, 1]``.  The pair list is the row-major non-zero set of ``outlet_topo`` exactly as
``(outlet_topo == 1).nonzero()`` produces it; it and its CSR/CSC index arrays depend on the
topology only and are cached per topology tensor, so the reference's per-call ``nonzero``
device-to-host sync (hbv_2_hourly.py:822-824) happens once.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _cabi as A
from .ops import _check_cuda, _stream, _timed

_TOPO_CACHE: dict = {}


class PairIndex:
    """Index arrays of one outlet topology (all int32, on the topology's device)."""

    def __init__(self, outlet_topo: torch.Tensor, areas: torch.Tensor):
        dev = outlet_topo.device
        idx = (outlet_topo == 1).nonzero(as_tuple=False)
        rows, cols = idx[:, 0].contiguous(), idx[:, 1].contiguous()
        self.n_gages, self.n_units = int(outlet_topo.shape[0]), int(outlet_topo.shape[1])
        self.n_pairs = int(rows.numel())
        self.pair_row = rows.to(torch.int32)
        self.pair_col = cols.to(torch.int32)
        z = torch.zeros(1, dtype=torch.int64, device=dev)
        self.gage_off = torch.cat([z, torch.bincount(rows, minlength=self.n_gages).cumsum(0)]).to(torch.int32)
        self.unit_off = torch.cat([z, torch.bincount(cols, minlength=self.n_units).cumsum(0)]).to(torch.int32)
        self.unit_perm = torch.argsort(cols, stable=True).to(torch.int32)
        a = areas.to(device=dev, dtype=torch.float32)
        denom = (outlet_topo.to(torch.float32) * a[None, :]).sum(dim=1).clamp(min=1e-6)  # :849
        self.inv_denom = (1.0 / denom).contiguous()
        self.areas = a.contiguous()


def pair_index(outlet_topo: torch.Tensor, areas: torch.Tensor) -> PairIndex:
    key = (outlet_topo.data_ptr(), tuple(outlet_topo.shape), outlet_topo._version,
           areas.data_ptr(), areas._version, str(outlet_topo.device))
    hit = _TOPO_CACHE.get(key)
    if hit is None:
        if len(_TOPO_CACHE) > 16:
            _TOPO_CACHE.clear()
        hit = _TOPO_CACHE[key] = PairIndex(outlet_topo, areas)
    return hit


def _desc(pi: PairIndex, T: int, lenF: int, lag_uh: bool, bounds) -> A.HbvPairDesc:
    d = A.HbvPairDesc()
    d.abi_version = A.ABI_VERSION
    d.T, d.n_pairs, d.n_units, d.n_gages, d.lenF = T, pi.n_pairs, pi.n_units, pi.n_gages, lenF
    d.lag_uh = int(lag_uh)
    (d.a_lo, d.a_hi), (d.b_lo, d.b_hi) = bounds[0], bounds[1]
    d.tau_lo, d.tau_hi = bounds[2] if len(bounds) > 2 else (0.0, 0.0)
    return d


class _PairRoute(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qs, par, pi: PairIndex, lenF, lag_uh, bounds):
        lib = A.load()
        dev = qs.device
        T = qs.shape[0]
        d = _desc(pi, T, lenF, lag_uh, bounds)
        M = min(T, lenF)
        uh = torch.empty((M, pi.n_pairs), device=dev, dtype=torch.float32)
        lag = torch.empty((T, pi.n_pairs), device=dev, dtype=torch.float32)
        out = torch.empty((T, pi.n_gages), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev), _timed('pair_route_fwd', dev):
            A.check(lib.hbv_b200_pair_route_fwd(
                C.byref(d), par.data_ptr(), qs.data_ptr(), pi.areas.data_ptr(), pi.pair_col.data_ptr(),
                pi.gage_off.data_ptr(), pi.inv_denom.data_ptr(), uh.data_ptr(), lag.data_ptr(),
                out.data_ptr(), _stream(dev)), 'pair_route_fwd')
        ctx.pi, ctx.args = pi, (lenF, lag_uh, bounds)
        ctx.save_for_backward(qs, par, uh)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = A.load()
        qs, par, uh = ctx.saved_tensors
        pi: PairIndex = ctx.pi
        lenF, lag_uh, bounds = ctx.args
        dev = qs.device
        T = qs.shape[0]
        d = _desc(pi, T, lenF, lag_uh, bounds)
        g_out = g_out.contiguous()
        nch = lib.hbv_b200_pair_chunks(T)
        g_lag = torch.empty((T, pi.n_pairs), device=dev, dtype=torch.float32)
        duh = torch.empty((uh.shape[0], nch, pi.n_pairs), device=dev, dtype=torch.float32)
        g_qs = torch.empty_like(qs)
        g_par = torch.empty_like(par)
        with torch.cuda.device(dev), _timed('pair_route_bwd', dev):
            A.check(lib.hbv_b200_pair_route_bwd(
                C.byref(d), par.data_ptr(), qs.data_ptr(), pi.areas.data_ptr(), pi.pair_col.data_ptr(),
                pi.pair_row.data_ptr(), pi.inv_denom.data_ptr(), pi.unit_off.data_ptr(),
                pi.unit_perm.data_ptr(), uh.data_ptr(), g_out.data_ptr(), g_lag.data_ptr(),
                duh.data_ptr(), g_qs.data_ptr(), g_par.data_ptr(), _stream(dev)), 'pair_route_bwd')
        return g_qs, g_par, None, None, None, None


def distr_routing(Qs, distr_params, outlet_topo, areas, lenF=72, lag_uh=True,
                  bounds=((0, 5.0), (0, 12.0), (0, 48.0))):
    """Qs [T, n_units, 1], distr_params [n_pairs, 3] in [0, 1] -> gage flow [T, n_gages, 1]."""
    _check_cuda(Qs, 'Qs')
    pi = pair_index(outlet_topo, areas)
    if distr_params.shape[0] != pi.n_pairs:
        raise ValueError(f'distr parameters have {distr_params.shape[0]} rows, topology has '
                         f'{pi.n_pairs} (gage, unit) pairs')
    qs = Qs.reshape(Qs.shape[0], -1).contiguous()
    par = distr_params.to(device=Qs.device, dtype=torch.float32).contiguous()
    if par.shape[1] < 3:
        par = torch.cat([par, par.new_zeros(par.shape[0], 3 - par.shape[1])], dim=1)
    out = _PairRoute.apply(qs, par, pi, lenF, lag_uh, tuple(tuple(b) for b in bounds))
    return out.unsqueeze(-1)
