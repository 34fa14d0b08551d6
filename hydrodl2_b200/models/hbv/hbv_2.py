"""HBV 2.0 — B200-native drop-in for ``hydrodl2/models/hbv/hbv_2.py:8-670``."""

from __future__ import annotations

import torch

from ... import _cabi as A
from ._split import _FLUX, SplitHbv


class Hbv_2(SplitHbv):
    """HBV 2.0: HBV 1.1p + lateral flux (parRT, parAC with `ac_all`), elevation-switched
    threshold temperature (`elev_all`), split (dynamic, static) parameters in [0, 1], no
    internal warm-up, routing off by default, per-step state series (hbv_2.py:324-390)."""

    _variant = A.VARIANT_HBV2
    _name = 'HBV 2.0'
    _lenF = 15
    _dt = 1.0

    def forward(self, x_dict: dict[str, torch.Tensor], parameters) -> dict[str, torch.Tensor]:
        x, dyn, sta, current, ngrid = self._prep(x_dict, parameters)
        attrs = self._attrs(x_dict)
        drop = self._draw_drop(ngrid)
        res = self._run(x, dyn, sta, current, attrs, drop, self.routing)
        self._store_states(res)
        if self.initialize:
            return {}
        return self._flux_dict(res, x)

    def _flux_from_run(self, res, x, outlet_topo=None, areas=None, distr_params_dict=None):
        return self._flux_dict(res, x)

    def _flux_dict(self, res, x) -> dict[str, torch.Tensor]:
        flux, routed = res['flux'], res['routed']
        out = {}
        if routed is not None:
            qs, q0, q1, q2 = routed
            bfi = res['bfi']
        else:  # hbv_2.py:618-626: un-routed means stand in for the routed series
            qs, q0, q1, q2 = flux[A.F_QSIM], flux[A.F_Q0], flux[A.F_Q1], flux[A.F_Q2]
            bfi = 100 * (q2.sum(0) / (qs.sum(0) + self.nearzero))
        out['streamflow'] = qs.unsqueeze(-1)
        out['srflow'] = q0.unsqueeze(-1)
        out['ssflow'] = q1.unsqueeze(-1)
        out['gwflow'] = q2.unsqueeze(-1)
        fl = dict(_FLUX)
        out['AET_hydro'] = flux[fl['AET_hydro']].unsqueeze(-1)
        out['PET_hydro'] = x[:, :, self.variables.index('pet')].unsqueeze(-1)
        for key, slot in _FLUX[1:]:
            out[key] = flux[slot].unsqueeze(-1)
        out['BFI'] = bfi
        if not self.warm_up_states:
            # hbv_2.py:666-669 slices with `pred_cutoff`, which the 2.0 models never set (it stays
            # 0): with warm_up_states=False the reference returns all T rows — so does this.
            for key in out:
                if key != 'BFI':
                    out[key] = out[key][self.pred_cutoff:, :, :]
        return out
