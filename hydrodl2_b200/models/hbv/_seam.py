"""Compatibility helpers for callers that drive the `_PBM` seam themselves.

The reference's ``forward`` unpacks, squashes and descales the network output into dictionaries
of ``[T, B, nmul]`` tensors before it calls ``_PBM`` (hbv.py:182-282, hbv_2.py:190-322,
hbv_2_hourly.py:254-374), and a subclass or an orchestrator such as ``Hbv_2_mts`` may call those
helpers and ``_PBM`` directly.  The models here never do — unpack / sigmoid / descale are fused
into the CUDA kernels — but they keep the helpers (same names, arguments and results, same
`torch.bernoulli` draw order) so such callers keep working, and expose ``routing_param_dict``
lazily from the last ``forward``.  Plain PyTorch tensor algebra on whatever device the inputs
live on; not on the product path.
"""

from __future__ import annotations

import torch


def _affine(t: torch.Tensor, bounds) -> torch.Tensor:
    lo, hi = bounds
    return t * (hi - lo) + lo          # core/calc/utils.py:24


class _LazyRouting:
    """`routing_param_dict` as the reference leaves it after forward (hbv.py:311-312), computed
    on first access from the normalised routing columns remembered by forward."""

    _routing_norm = None
    _routing_dict = None

    @property
    def routing_param_dict(self):
        if self._routing_dict is None and self._routing_norm is not None:
            self._routing_dict = self._descale_route_parameters(self._routing_norm())
        if self._routing_dict is None:
            raise AttributeError('routing_param_dict: set it (or run forward with routing=True) before '
                                 'calling _PBM with routing enabled')
        return self._routing_dict

    @routing_param_dict.setter
    def routing_param_dict(self, value):
        self._routing_dict = value
        self._routing_norm = None

    def _remember_routing(self, fn) -> None:
        self._routing_norm, self._routing_dict = fn, None

    def _descale_route_parameters(self, routing_params: torch.Tensor) -> dict:
        """[B, 2] in [0, 1] -> {'route_a': [B], 'route_b': [B]} in physical range."""
        return {name: _affine(routing_params[:, i], self.routing_parameter_bounds[name])
                for i, name in enumerate(self.routing_parameter_bounds.keys())}


class PackedSeam(_LazyRouting):
    """hbv.py:182-282 / hbv_1_1p.py for the packed raw-parameter models."""

    def _unpack_parameters(self, parameters: torch.Tensor):
        n = len(self.parameter_bounds)
        T, B = parameters.shape[0], parameters.shape[1]
        phy = torch.sigmoid(parameters[:, :, :n * self.nmul]).view(T, B, n, self.nmul)
        routing = torch.sigmoid(parameters[-1, :, n * self.nmul:]) if self.routing else None
        return phy, routing

    def _descale_phy_parameters(self, phy_params: torch.Tensor, dy_list: list) -> dict:
        T, B = phy_params.shape[0], phy_params.shape[1]
        pmat = torch.ones([1, B, 1]) * self.dy_drop
        out = {}
        for i, name in enumerate(self.parameter_bounds.keys()):
            static = phy_params[-1, :, i, :].unsqueeze(0).expand(T, B, self.nmul)
            if name in dy_list:
                # one CPU draw per dynamic parameter, in bounds order (hbv.py:240-246)
                keep_static = torch.bernoulli(pmat).detach_().to(phy_params.device)
                value = phy_params[:, :, i, :] * (1 - keep_static) + static * keep_static
            else:
                value = static
            out[name] = _affine(value, self.parameter_bounds[name])
        return out


class SplitSeam(_LazyRouting):
    """hbv_2.py:190-322 / hbv_2_hourly.py:254-374 for the (dynamic, static[, distr]) tuple form."""

    def _unpack_parameters(self, parameters):
        n_dy = len(self.dynamic_params)
        n_sta = len(self.parameter_bounds) - n_dy
        dyn, sta = parameters[0], parameters[1]
        phy_dy = dyn.view(dyn.shape[0], dyn.shape[1], n_dy, self.nmul)
        phy_sta = sta[:, :n_sta * self.nmul].view(sta.shape[0], n_sta, self.nmul)
        routing = sta[:, n_sta * self.nmul:] if self.routing else None
        if hasattr(self, 'distr_parameter_bounds'):
            return phy_dy, phy_sta, routing, (parameters[2] if len(parameters) > 2 else None)
        return phy_dy, phy_sta, routing

    def _descale_phy_dy_parameters(self, phy_dy_params: torch.Tensor, dy_list: list) -> dict:
        T, B = phy_dy_params.shape[0], phy_dy_params.shape[1]
        pmat = torch.ones([1, B, 1]) * self.dy_drop
        out = {}
        for i, name in enumerate(dy_list):
            static = phy_dy_params[-1, :, i, :].unsqueeze(0).expand(T, B, self.nmul)
            keep_static = torch.bernoulli(pmat).detach_().to(phy_dy_params.device)
            value = phy_dy_params[:, :, i, :] * (1 - keep_static) + static * keep_static
            out[name] = _affine(value, self.parameter_bounds[name])
        return out

    def _descale_phy_stat_parameters(self, phy_stat_params: torch.Tensor, stat_list: list) -> dict:
        return {name: _affine(phy_stat_params[:, i, :], self.parameter_bounds[name])
                for i, name in enumerate(stat_list)}
