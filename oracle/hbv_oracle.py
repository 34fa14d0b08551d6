"""CPU restatement of the hydrodl2 HBV recurrence + gamma-UH routing.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hydrodl2_b200/`` may import this
module: it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
is the CUDA library and fails loudly without it.

The functions restate, op for op and in the reference's evaluation order, the
PyTorch arithmetic of (paths relative to ``/root/reference/src/hydrodl2``):

* ``models/hbv/hbv.py:182-256``      sigmoid + static/dynamic descaling
* ``models/hbv/hbv.py:423-505``      HBV 1.0 step
* ``models/hbv/hbv_1_1p.py:422-516`` HBV 1.1p step (BETAET always, capillary)
* ``models/hbv/hbv_2.py:464-575``    HBV 2.0 step (elevation TT switch, lateral
                                      flux, per-step state series)
* ``models/hbv/hbv_2_hourly.py:527-675`` hourly step (dt algebra, guard rails,
                                      Hortonian infiltration)
* ``models/hbv/hbv_2_hourly.py:800-897`` distributed (gage, unit) pair routing
* ``core/calc/uh_routing.py:5-57``   ``uh_gamma`` / ``uh_conv``
* ``core/calc/utils.py:9-24``        ``change_param_range``

They are written as pure functions over explicit tensors (no nn.Module, no
hidden state) and run in any float dtype: float32 reproduces the reference's
CPU results bit for bit (pinned by ``tests/golden/*.npz`` which were produced
by importing the unmodified reference, see ``tests/golden/make_golden.py``),
float64 is the arbiter used to decide whether a float32 difference matters.
Gradients of the oracle come from PyTorch autograd over this restatement, i.e.
with exactly the sub-gradient conventions the reference gets (clamp inclusive,
``min`` ties split 1/2, mask casts carry no gradient).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# Variant tables (hbv.py:88-105, hbv_1_1p.py:87-106, hbv_2.py:90-111,
# hbv_2_hourly.py:91-124).  Order is part of the contract: column i*nmul+j.
# --------------------------------------------------------------------------
_BASE_BOUNDS = [
    ('parBETA', (1.0, 6.0)),
    ('parFC', (50, 1000)),
    ('parK0', (0.05, 0.9)),
    ('parK1', (0.01, 0.5)),
    ('parK2', (0.001, 0.2)),
    ('parLP', (0.2, 1)),
    ('parPERC', (0, 10)),
    ('parUZL', (0, 100)),
    ('parTT', (-2.5, 2.5)),
    ('parCFMAX', (0.5, 10)),
    ('parCFR', (0, 0.1)),
    ('parCWH', (0, 0.2)),
]
_BETAET = ('parBETAET', (0.3, 5))
_C = ('parC', (0, 1))
_RT = ('parRT', (0, 20))
_AC = ('parAC', (0, 2500))


@dataclass
class Variant:
    """Static description of one HBV variant."""

    name: str
    bounds: dict
    route_bounds: dict
    betaet: bool = False      # apply ** parBETAET
    capillary: bool = False   # 1.1p capillary rise
    lateral: bool = False     # 2.0 elevation TT switch + lateral flux
    hourly: bool = False      # dt algebra, guard rails, infiltration
    dt: float = 1.0
    lenF: int = 15
    sigmoid: bool = True      # raw parameters need a sigmoid (1.0 / 1.1p)
    distr_bounds: dict = field(default_factory=dict)


def variant(name: str, dynamic_params=()) -> Variant:
    """Build the variant table the reference's ``__init__`` would build."""
    if name == 'hbv':
        b = dict(_BASE_BOUNDS)
        betaet = 'parBETAET' in dynamic_params  # hbv.py:124-125
        if betaet:
            b['parBETAET'] = _BETAET[1]
        return Variant('hbv', b, {'route_a': (0, 2.9), 'route_b': (0, 6.5)},
                       betaet=betaet)
    if name == 'hbv_1_1p':
        b = dict(_BASE_BOUNDS + [_BETAET, _C])
        return Variant('hbv_1_1p', b, {'route_a': (0, 2.9), 'route_b': (0, 6.5)},
                       betaet=True, capillary=True)
    if name == 'hbv_2':
        b = dict(_BASE_BOUNDS + [_BETAET, _C, _RT, _AC])
        return Variant('hbv_2', b, {'route_a': (0, 2.9), 'route_b': (0, 6.5)},
                       betaet=True, capillary=True, lateral=True, sigmoid=False)
    if name == 'hbv_2_hourly':
        dt = 1.0 / 24
        b = dict(_BASE_BOUNDS + [_BETAET, _C, _RT, _AC,
                                 ('parF0', (5.0 / dt, 120.0 / dt)),
                                 ('parFMIN', (0.0, 1.0)),
                                 ('parALPHA', (0.5, 5.0))])
        return Variant('hbv_2_hourly', b, {'route_a': (0, 5.0), 'route_b': (0, 12.0)},
                       betaet=True, capillary=True, lateral=True, hourly=True,
                       dt=dt, lenF=72, sigmoid=False,
                       distr_bounds={'route_a': (0, 5.0), 'route_b': (0, 12.0),
                                     'route_tau': (0, 48.0)})
    raise ValueError(name)


def change_param_range(p, bounds):
    """core/calc/utils.py:24."""
    return p * (bounds[1] - bounds[0]) + bounds[0]


# --------------------------------------------------------------------------
# Routing (core/calc/uh_routing.py)
# --------------------------------------------------------------------------
def uh_gamma(a, b, lenF=10):
    """Normalised gamma-pdf unit hydrograph, uh_routing.py:5-22.

    a, b: [T, B, V] (only the first lenF rows are read) -> [lenF', B, V].
    """
    m = a.shape
    lenF = min(a.shape[0], lenF)
    aa = F.relu(a[0:lenF]).view(lenF, m[1], m[2]) + 0.1
    theta = F.relu(b[0:lenF]).view(lenF, m[1], m[2]) + 0.5
    t = torch.arange(0.5, lenF * 1.0).view(lenF, 1, 1).repeat(1, m[1], m[2])
    t = t.to(device=aa.device, dtype=aa.dtype)
    denom = (aa.lgamma().exp()) * (theta ** aa)
    mid = t ** (aa - 1)
    right = torch.exp(-t / theta)
    w = 1 / denom * mid * right
    return w / w.sum(0)


def uh_conv(x, UH):
    """Causal per-basin convolution, uh_routing.py:25-57 (viewmode 1).

    x: [B, 1, T], UH: [B, 1, m] -> y[b, 0, t] = sum_k UH[b,0,k] x[b,0,t-k].
    """
    nb, _, nt = x.shape
    m = UH.shape[-1]
    y = F.conv1d(x.view(1, nb, nt), torch.flip(UH.view(nb, 1, m), [2]),
                 groups=nb, padding=m - 1)
    if m > 1:
        y = y[:, :, 0:-(m - 1)]
    return y.view(x.shape)


def uh_conv_direct(x, UH):
    """Same sum written as an explicit shift-and-add (definition check)."""
    nb, _, nt = x.shape
    m = UH.shape[-1]
    y = torch.zeros_like(x)
    for k in range(min(m, nt)):
        y[:, :, k:] = y[:, :, k:] + UH[:, :, k:k + 1] * x[:, :, :nt - k]
    return y


def frac_shift1d(w, tau):
    """hbv_2_hourly.py:858-897: y[t] = (1-f) w[t-k] + f w[t-k-1], zero fill."""
    T, B, V = w.shape
    tau = tau.view(1, B, V).to(w.dtype)
    k = torch.floor(tau)
    f = tau - k
    t = torch.arange(T, device=w.device, dtype=w.dtype).view(T, 1, 1)
    i0 = t - k
    i1 = t - (k + 1)
    w0 = torch.gather(w, 0, i0.clamp(0, T - 1).long())
    w1 = torch.gather(w, 0, i1.clamp(0, T - 1).long())
    w0 = w0 * ((i0 >= 0) & (i0 <= T - 1)).to(w.dtype)
    w1 = w1 * ((i1 >= 0) & (i1 <= T - 1)).to(w.dtype)
    return (1.0 - f) * w0 + f * w1


def distr_routing(Qs, distr, outlet_topo, areas, lenF=72, lag_uh=True):
    """hbv_2_hourly.py:800-855.  Qs [T, n_units, 1] -> [T, n_gages, 1]."""
    nsteps = Qs.size(0)
    Qw = Qs * areas[None, :, None]
    idx = (outlet_topo == 1).nonzero(as_tuple=False)
    rows, cols = idx[:, 0].long(), idx[:, 1].long()
    Qp = Qw[:, cols, :]
    UH = uh_gamma(distr['route_a'].repeat(nsteps, 1).unsqueeze(-1),
                  distr['route_b'].repeat(nsteps, 1).unsqueeze(-1), lenF=lenF)
    if lag_uh:
        UH = frac_shift1d(UH, distr['route_tau'])
    rf = Qp.permute(1, 2, 0).contiguous()
    UH = UH.permute(1, 2, 0).contiguous()
    lag = uh_conv(rf, UH).squeeze(1).contiguous()
    n_gages = int(outlet_topo.shape[0])
    out = torch.zeros(n_gages, lag.shape[1], dtype=lag.dtype, device=lag.device)
    out = out.scatter_add(0, rows.view(-1, 1).expand(-1, lag.shape[1]), lag)
    denom = (outlet_topo * areas[None, :]).sum(dim=1).unsqueeze(1).clamp(min=1e-6)
    return (out / denom).T.unsqueeze(-1)


# --------------------------------------------------------------------------
# One time step
# --------------------------------------------------------------------------
def hbv_step(v: Variant, S, p, Pm, Tm, PETm, Ac=None, Elev=None, nearzero=1e-5):
    """One step of the recurrence for every (basin, component).

    S = (SNOWPACK, MELTWATER, SM, SUZ, SLZ), each [B, nmul];  p: dict of
    descaled parameters [B, nmul];  Pm/Tm/PETm: [B, nmul] (hourly: already
    divided by dt as hbv_2_hourly.py:485-487 does).  Returns (S', fluxes).
    """
    SNOWPACK, MELTWATER, SM, SUZ, SLZ = S
    f32 = Pm.dtype
    dt = v.dt

    if v.hourly:  # hbv_2_hourly.py:528-533
        SNOWPACK = torch.clamp(SNOWPACK, min=0.0)
        MELTWATER = torch.clamp(MELTWATER, min=0.0)
        SM = torch.clamp(SM, min=nearzero)
        SUZ = torch.clamp(SUZ, min=nearzero)
        SLZ = torch.clamp(SLZ, min=nearzero)

    if v.lateral:  # hbv_2.py:473-477
        TT = (Elev >= 2000).type(f32) * 4.0 + (Elev < 2000).type(f32) * p['parTT']
    else:
        TT = p['parTT']
    RAIN = torch.mul(Pm, (Tm >= TT).type(f32))
    SNOW = torch.mul(Pm, (Tm < TT).type(f32))

    # Snow (hbv.py:438-459; hourly :551-572)
    if v.hourly:
        SNOWPACK = SNOWPACK + SNOW * dt
        melt = torch.clamp(p['parCFMAX'] * (Tm - TT), min=0.0)
        melt = torch.min(melt * dt, SNOWPACK)
    else:
        SNOWPACK = SNOWPACK + SNOW
        melt = torch.clamp(p['parCFMAX'] * (Tm - TT), min=0.0)
        melt = torch.min(melt, SNOWPACK)
    MELTWATER = MELTWATER + melt
    SNOWPACK = SNOWPACK - melt
    refreezing = torch.clamp(p['parCFR'] * p['parCFMAX'] * (TT - Tm), min=0.0)
    if v.hourly:
        refreezing = torch.min(refreezing * dt, MELTWATER)
    else:
        refreezing = torch.min(refreezing, MELTWATER)
    SNOWPACK = SNOWPACK + refreezing
    MELTWATER = MELTWATER - refreezing
    if v.hourly:
        tosoil = torch.clamp((MELTWATER - (p['parCWH'] * SNOWPACK)) / dt, min=0.0)
        MELTWATER = MELTWATER - tosoil * dt
    else:
        tosoil = torch.clamp(MELTWATER - (p['parCWH'] * SNOWPACK), min=0.0)
        MELTWATER = MELTWATER - tosoil

    # Soil (hbv.py:462-480; hourly :575-617)
    IE = None
    if v.hourly:
        W = RAIN + tosoil
        s = torch.clamp(SM / p['parFC'], 0.0, 1.0 - 0.01)
        fmin = p['parFMIN'] * p['parF0']
        fcap = fmin + (p['parF0'] - fmin) * torch.pow(1.0 - s, p['parALPHA'])
        infiltration = torch.minimum(W, fcap)
        IE = torch.clamp(W - fcap, min=0.0)
        soil_wetness = torch.clamp((SM / p['parFC']) ** p['parBETA'], 0.0, 1.0)
        recharge = infiltration * soil_wetness
        SM = SM + (infiltration - recharge) * dt
        excess = torch.clamp((SM - p['parFC']) / dt, min=0.0)
        SM = SM - excess * dt
    else:
        soil_wetness = (SM / p['parFC']) ** p['parBETA']
        soil_wetness = torch.clamp(soil_wetness, min=0.0, max=1.0)
        recharge = (RAIN + tosoil) * soil_wetness
        SM = SM + RAIN + tosoil - recharge
        excess = torch.clamp(SM - p['parFC'], min=0.0)
        SM = SM - excess
    evapfactor = SM / (p['parLP'] * p['parFC'])
    if v.betaet:
        evapfactor = evapfactor ** p['parBETAET']
    evapfactor = torch.clamp(evapfactor, min=0.0, max=1.0)
    ETact = PETm * evapfactor
    if v.hourly:
        ETact = torch.min(SM, ETact * dt) / dt
        SM = torch.clamp(SM - ETact * dt, min=nearzero)
    else:
        ETact = torch.min(SM, ETact)
        SM = torch.clamp(SM - ETact, min=nearzero)

    capillary = None
    if v.capillary:  # hbv_1_1p.py:482-490; hourly :620-633
        cap = p['parC'] * SLZ * (1.0 - torch.clamp(SM / p['parFC'], max=1.0))
        if v.hourly:
            capillary = torch.min(SLZ, cap * dt) / dt
            SM = torch.clamp(SM + capillary * dt, min=nearzero)
            SLZ = torch.clamp(SLZ - capillary * dt, min=nearzero)
        else:
            capillary = torch.min(SLZ, cap)
            SM = torch.clamp(SM + capillary, min=nearzero)
            SLZ = torch.clamp(SLZ - capillary, min=nearzero)

    # Groundwater boxes (hbv.py:483-492; hbv_2.py:535-553; hourly :636-655)
    if v.hourly:
        SUZ = SUZ + (recharge + excess) * dt
        PERC = torch.min(SUZ, p['parPERC'] * dt) / dt
        SUZ = SUZ - PERC * dt
        Q0 = p['parK0'] * torch.clamp(SUZ - p['parUZL'], min=0.0)
        SUZ = SUZ - Q0 * dt
        Q1 = p['parK1'] * SUZ
        SUZ = SUZ - Q1 * dt
        SLZ = SLZ + PERC * dt
    else:
        SUZ = SUZ + recharge + excess
        PERC = torch.min(SUZ, p['parPERC'])
        SUZ = SUZ - PERC
        Q0 = p['parK0'] * torch.clamp(SUZ - p['parUZL'], min=0.0)
        SUZ = SUZ - Q0
        Q1 = p['parK1'] * SUZ
        SUZ = SUZ - Q1
        SLZ = SLZ + PERC
    if v.lateral:
        LF = torch.clamp((Ac - p['parAC']) / 1000, min=-1, max=1) * p['parRT'] * (
            Ac < 2500
        ) + torch.exp(torch.clamp(-(Ac - 2500) / 50, min=-10.0, max=0.0)) * p[
            'parRT'
        ] * (Ac >= 2500)
        SLZ = torch.clamp(SLZ + (LF * dt if v.hourly else LF), min=0.0)
    Q2 = p['parK2'] * SLZ
    SLZ = SLZ - (Q2 * dt if v.hourly else Q2)

    Qsim = Q0 + Q1 + Q2
    if v.hourly:
        Qsim = Qsim + IE
    flux = {
        'Qsim': Qsim, 'Q0': Q0, 'Q1': Q1, 'Q2': Q2, 'AET': ETact,
        'SWE': SNOWPACK, 'recharge': recharge, 'excs': excess,
        'evapfactor': evapfactor, 'tosoil': tosoil, 'PERC': PERC,
    }
    if v.capillary:
        flux['capillary'] = capillary
    return (SNOWPACK, MELTWATER, SM, SUZ, SLZ), flux


# --------------------------------------------------------------------------
# Parameter handling
# --------------------------------------------------------------------------
def descale_packed(v: Variant, phy01, dy_list, dy_drop, rng_masks=None):
    """hbv.py:217-256.  phy01: [T', B, n, nmul] in [0,1] -> dict name->[T',B,nmul].

    Static value = LAST row of the slice handed in (hbv.py:242).  One
    ``torch.bernoulli`` draw per dynamic parameter, in bounds order, on CPU —
    the same RNG consumption as the reference.  ``rng_masks`` (name -> [B])
    overrides the draw.
    """
    nsteps, ngrid = phy01.shape[0], phy01.shape[1]
    out = {}
    pmat = torch.ones([1, ngrid, 1]) * dy_drop
    for i, name in enumerate(v.bounds.keys()):
        sta = phy01[-1, :, i, :].unsqueeze(0).expand(nsteps, -1, -1)
        if name in dy_list:
            if rng_masks is not None:
                dr = rng_masks[name].view(1, ngrid, 1).to(phy01)
            else:
                dr = torch.bernoulli(pmat).detach_().to(phy01)
            com = phy01[:, :, i, :] * (1 - dr) + sta * dr
            out[name] = change_param_range(com, v.bounds[name])
        else:
            out[name] = change_param_range(sta, v.bounds[name])
    return out


def descale_split(v: Variant, dyn01, sta01, dy_list, dy_drop, rng_masks=None):
    """hbv_2.py:232-290.  dyn01 [T,B,n_dy,nmul], sta01 [B,n_sta,nmul]."""
    nsteps, ngrid = dyn01.shape[0], dyn01.shape[1]
    dyn = {}
    pmat = torch.ones([1, ngrid, 1]) * dy_drop
    for i, name in enumerate(dy_list):
        sta = dyn01[-1, :, i, :].unsqueeze(0).expand(nsteps, -1, -1)
        if rng_masks is not None:
            dr = rng_masks[name].view(1, ngrid, 1).to(dyn01)
        else:
            dr = torch.bernoulli(pmat).detach_().to(dyn01)
        com = dyn01[:, :, i, :] * (1 - dr) + sta * dr
        dyn[name] = change_param_range(com, v.bounds[name])
    stat = {}
    stat_list = [n for n in v.bounds.keys() if n not in dy_list]
    for i, name in enumerate(stat_list):
        stat[name] = change_param_range(sta01[:, i, :], v.bounds[name])
    return dyn, stat


# --------------------------------------------------------------------------
# The recurrence over time + aggregation + routing (the `_PBM` equivalent)
# --------------------------------------------------------------------------
def run_pbm(v: Variant, forcing, states, dyn, stat=None, *, Ac=None, Elev=None,
            nmul=16, nearzero=1e-5, muwts=None, routing=True, route=None,
            initialize=False, variables=('prcp', 'tmean', 'pet'),
            keep_state_series=False):
    """hbv.py:363-596 / hbv_2.py:392-660 / hbv_2_hourly.py:451-760.

    dyn:  name -> [T, B, nmul] (time-varying or time-replicated)
    stat: name -> [B, nmul]    (hbv_2 family static parameters), may be None
    Returns (flux_dict, final_states, state_series or None).
    """
    P = forcing[:, :, variables.index('prcp')]
    Tt = forcing[:, :, variables.index('tmean')]
    PET = forcing[:, :, variables.index('pet')]
    if v.hourly:
        P = P / v.dt
        PET = PET / v.dt
    nsteps, ngrid = P.shape
    Pm = P.unsqueeze(2).repeat(1, 1, nmul)
    Tm = Tt.unsqueeze(2).repeat(1, 1, nmul)
    PETm = PET.unsqueeze(-1).repeat(1, 1, nmul)

    S = tuple(states)
    series = {}
    sseries = [[] for _ in range(5)]
    for t in range(nsteps):
        p = {k: val[t] for k, val in dyn.items()}
        if stat:
            p.update(stat)
        S, fl = hbv_step(v, S, p, Pm[t], Tm[t], PETm[t], Ac, Elev, nearzero)
        if not initialize:
            for k, val in fl.items():
                series.setdefault(k, []).append(val)
        if keep_state_series:
            for i in range(5):
                sseries[i].append(S[i])
    state_series = tuple(torch.stack(s) for s in sseries) if keep_state_series else None
    if initialize:
        return {}, S, state_series

    mu = {k: torch.stack(val) for k, val in series.items()}  # [T,B,nmul]
    if muwts is None:
        Qsimavg = mu['Qsim'].mean(-1)
    else:
        Qsimavg = (mu['Qsim'] * muwts).sum(-1)

    def mean(k):
        return mu[k].mean(-1, keepdim=True)

    if routing:
        UH = uh_gamma(route['route_a'].repeat(nsteps, 1).unsqueeze(-1),
                      route['route_b'].repeat(nsteps, 1).unsqueeze(-1), lenF=v.lenF)
        UH = UH.permute(1, 2, 0)
        Qs = uh_conv(Qsimavg.unsqueeze(-1).permute(1, 2, 0), UH).permute(2, 0, 1)
        if not v.hourly:
            Q0r = uh_conv(mean('Q0').permute(1, 2, 0), UH).permute(2, 0, 1)
            Q1r = uh_conv(mean('Q1').permute(1, 2, 0), UH).permute(2, 0, 1)
            Q2r = uh_conv(mean('Q2').permute(1, 2, 0), UH).permute(2, 0, 1)
    else:
        Qs = Qsimavg.unsqueeze(-1)
        if v.lateral and not v.hourly:  # hbv_2.py:620-626
            Q0r, Q1r, Q2r = mean('Q0'), mean('Q1'), mean('Q2')
        else:
            Q0r = Q1r = Q2r = None

    if v.hourly:  # hbv_2_hourly.py:740-741
        return {'Qs': Qs * v.dt}, S, state_series

    BFI = 100 * (torch.sum(Q2r, dim=0) / (torch.sum(Qs, dim=0) + nearzero))[:, 0]
    out = {
        'streamflow': Qs, 'srflow': Q0r, 'ssflow': Q1r, 'gwflow': Q2r,
        'AET_hydro': mean('AET'), 'PET_hydro': PETm.mean(-1, keepdim=True),
        'SWE': mean('SWE'), 'streamflow_no_rout': Qsimavg.unsqueeze(2),
        'srflow_no_rout': mean('Q0'), 'ssflow_no_rout': mean('Q1'),
        'gwflow_no_rout': mean('Q2'), 'recharge': mean('recharge'),
        'excs': mean('excs'), 'evapfactor': mean('evapfactor'),
        'tosoil': mean('tosoil'), 'percolation': mean('PERC'),
    }
    if v.capillary:
        out['capillary'] = mean('capillary')
    out['BFI'] = BFI
    return out, S, state_series


# --------------------------------------------------------------------------
# Model-level entry points (the `forward` equivalents)
# --------------------------------------------------------------------------
def init_states(ngrid, nmul, dtype=torch.float32, device='cpu'):
    """hbv.py:128-136."""
    return tuple(torch.full((ngrid, nmul), 0.001, dtype=dtype, device=device)
                 for _ in range(5))


def forward_packed(name, x_phy, parameters, *, nmul=16, warm_up=0,
                   dynamic_params=(), dy_drop=0.0, warm_up_states=True,
                   nearzero=1e-5, muwts=None, states=None, dtype=None,
                   rng_masks=None):
    """``Hbv.forward`` / ``Hbv_1_1p.forward`` (hbv.py:284-361).

    Returns (flux_dict, final_states).  ``dtype=torch.float64`` runs the whole
    chain (sigmoid included) in double precision.
    """
    v = variant(name, dynamic_params)
    if dtype is not None:
        x_phy = x_phy.to(dtype)
        parameters = parameters.to(dtype)
    n = len(v.bounds)
    T, B = parameters.shape[0], parameters.shape[1]
    phy01 = torch.sigmoid(parameters[:, :, :n * nmul]).view(T, B, n, nmul)
    r01 = torch.sigmoid(parameters[-1, :, n * nmul:])
    route = {k: change_param_range(r01[:, i], bd)
             for i, (k, bd) in enumerate(v.route_bounds.items())}
    pred_cutoff = 0
    if not warm_up_states:
        pred_cutoff, warm_up = warm_up, 0
    S = states if states is not None else init_states(B, nmul, x_phy.dtype)
    if warm_up > 0:
        with torch.no_grad():
            pw = descale_packed(v, phy01[:warm_up], [], dy_drop)
            _, S, _ = run_pbm(v, x_phy[:warm_up], S, pw, nmul=nmul,
                              nearzero=nearzero, initialize=True, routing=False)
    pm = descale_packed(v, phy01[warm_up:], list(dynamic_params), dy_drop, rng_masks)
    out, S, _ = run_pbm(v, x_phy[warm_up:], S, pm, nmul=nmul, nearzero=nearzero,
                        muwts=muwts, routing=True, route=route)
    if pred_cutoff:
        out = {k: (val if k == 'BFI' else val[pred_cutoff:]) for k, val in out.items()}
    return out, S


def forward_split(name, x_dict, parameters, *, nmul=16, dynamic_params=(),
                  dy_drop=0.0, nearzero=1e-5, routing=False, states=None,
                  dtype=None, rng_masks=None, use_distr_routing=True):
    """``Hbv_2.forward`` (hbv_2.py:324-390) / ``Hbv_2_hourly.forward``
    (hbv_2_hourly.py:376-449).  Returns (flux_dict, state_series)."""
    v = variant(name, dynamic_params)
    x = x_dict['x_phy']
    conv = (lambda z: z.to(dtype)) if dtype is not None else (lambda z: z)
    x = conv(x)
    p0, p1 = conv(parameters[0]), conv(parameters[1])
    n, ndy = len(v.bounds), len(dynamic_params)
    nsta = n - ndy
    T, B = x.shape[0], x.shape[1]
    Ac = conv(x_dict['ac_all']).unsqueeze(-1).repeat(1, nmul)
    Elev = conv(x_dict['elev_all']).unsqueeze(-1).repeat(1, nmul)
    dyn01 = p0.view(p0.shape[0], p0.shape[1], ndy, nmul)
    sta01 = p1[:, :nsta * nmul].view(p1.shape[0], nsta, nmul)
    route = None
    if routing:
        r = p1[:, nsta * nmul:]
        route = {k: change_param_range(r[:, i], bd)
                 for i, (k, bd) in enumerate(v.route_bounds.items())}
    dyn, stat = descale_split(v, dyn01, sta01, list(dynamic_params), dy_drop, rng_masks)
    S = states if states is not None else init_states(B, nmul, x.dtype)
    out, _, series = run_pbm(v, x, S, dyn, stat, Ac=Ac, Elev=Elev, nmul=nmul,
                             nearzero=nearzero, muwts=x_dict.get('muwts'),
                             routing=routing, route=route, keep_state_series=True)
    if v.hourly and use_distr_routing:
        d = conv(parameters[2])
        distr = {k: change_param_range(d[:, i], bd)
                 for i, (k, bd) in enumerate(v.distr_bounds.items())}
        out['streamflow'] = distr_routing(out['Qs'], distr, conv(x_dict['outlet_topo']),
                                          conv(x_dict['areas']), lenF=v.lenF)
    return out, series


def forward_mts(x_dict, parameters, *, nmul=16, low_dynamic=(), high_dynamic=(), nearzero=1e-5,
                dtype=None):
    """``Hbv_2_mts._forward`` (hbv_2_mts.py:100-174), un-chunked training path: daily Hbv_2
    warm-up (states detached) -> identity state transfer -> param_transfer (hbv_2_mts.py:292-341)
    -> Hbv_2_hourly._PBM without distributed routing.  Returns ({'Qs'}, hourly state series)."""
    conv = (lambda z: z.to(dtype)) if dtype is not None else (lambda z: z)
    (lo_dyn, lo_sta), hi = parameters
    hi_dyn, hi_sta = hi[0], hi[1]
    xl = {'x_phy': x_dict['x_phy_low_freq'], 'ac_all': x_dict['ac_all'], 'elev_all': x_dict['elev_all']}
    _, series = forward_split('hbv_2', xl, [lo_dyn, lo_sta], nmul=nmul, dynamic_params=low_dynamic,
                              nearzero=nearzero, dtype=dtype)
    states = tuple(s[-1].detach() for s in series)
    vl, vh = variant('hbv_2', low_dynamic), variant('hbv_2_hourly', high_dynamic)
    hi_names = [n for n in vh.bounds if n not in high_dynamic]
    lo_names = [n for n in vl.bounds if n not in low_dynamic]
    lo3 = conv(lo_sta)[:, :len(lo_names) * nmul].view(-1, len(lo_names), nmul)
    hi3 = conv(hi_sta)[:, :len(hi_names) * nmul].view(-1, len(hi_names), nmul)
    extra = [i for i, n in enumerate(hi_names) if n not in lo_names]
    merged = torch.cat([lo3, hi3[:, extra]], dim=1)
    x = conv(x_dict['x_phy_high_freq'])
    T, B = x.shape[0], x.shape[1]
    dyn01 = conv(hi_dyn).view(T, B, len(high_dynamic), nmul)
    dyn, stat = descale_split(vh, dyn01, merged, list(high_dynamic), 0.0)
    Ac = conv(x_dict['ac_all']).unsqueeze(-1).repeat(1, nmul)
    Elev = conv(x_dict['elev_all']).unsqueeze(-1).repeat(1, nmul)
    out, _, hs = run_pbm(vh, x, states, dyn, stat, Ac=Ac, Elev=Elev, nmul=nmul, nearzero=nearzero,
                         routing=False, keep_state_series=True)
    return out, hs


# --------------------------------------------------------------------------
# Synthetic inputs — shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------
def synthetic_forcing(T, B, seed=20261017, hourly=False):
    """[T, B, 3] = (prcp, tmean, pet); seasonal temperature crossing TT."""
    g = torch.Generator().manual_seed(seed)
    import math
    steps_per_day = 24 if hourly else 1
    nd = (T + steps_per_day - 1) // steps_per_day
    d = torch.arange(nd, dtype=torch.float32).view(nd, 1)
    ob = torch.rand(1, B, generator=g) * 16 - 8
    season = torch.sin(2 * math.pi * (d - 110) / 365)
    tmean = 5 + 12 * season + ob + 4 * torch.randn(nd, B, generator=g)
    prcp = 5 * torch.relu(torch.randn(nd, B, generator=g))
    pet = torch.relu(2 + 2 * season) + 0.5 * torch.rand(nd, B, generator=g)
    x = torch.stack([prcp, tmean, pet], dim=-1)
    if hourly:
        x = x.repeat_interleave(24, dim=0)[:T].clone()
        x[:, :, 0] /= 24
        x[:, :, 2] /= 24
    return x.contiguous()
