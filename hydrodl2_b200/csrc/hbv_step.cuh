// hbv_step.cuh — one HBV time step for one (basin, component) lane, forward and adjoint.
//
// Forward arithmetic follows, in the reference's evaluation order,
//   models/hbv/hbv.py:428-492          (HBV 1.0)
//   models/hbv/hbv_1_1p.py:427-503     (+ BETAET always, capillary rise)
//   models/hbv/hbv_2.py:471-553        (+ elevation TT switch, lateral flux)
//   models/hbv/hbv_2_hourly.py:527-655 (+ dt algebra, guard rails, Hortonian infiltration)
// (paths relative to /root/reference/src/hydrodl2).  The adjoint is derived by hand and
// reproduces PyTorch's sub-gradient conventions (SURVEY.md §8 a-notes 2):
//   clamp(x, lo, hi): 1 where lo <= x <= hi (inclusive); min(a, b): all to the smaller,
//   1/2 : 1/2 on ties; mask casts carry no gradient; a**b: d/da = b a^b / a, d/db = a^b ln a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hbv_b200.h"

namespace hbv {

template <int VAR> struct Traits;
template <> struct Traits<HBV_VARIANT_HBV> {
    static constexpr bool CAP = false, LAT = false, HOURLY = false;
    static constexpr int NPAR = 13, NFLUX = 11;
};
template <> struct Traits<HBV_VARIANT_HBV11P> {
    static constexpr bool CAP = true, LAT = false, HOURLY = false;
    static constexpr int NPAR = 14, NFLUX = 12;
};
template <> struct Traits<HBV_VARIANT_HBV2> {
    static constexpr bool CAP = true, LAT = true, HOURLY = false;
    static constexpr int NPAR = 16, NFLUX = 12;
};
template <> struct Traits<HBV_VARIANT_HOURLY> {
    static constexpr bool CAP = true, LAT = true, HOURLY = true;
    static constexpr int NPAR = 19, NFLUX = 12;
};

// Per-lane constants that do not change over time.
struct LaneConst {
    float Ac, Elev;     // hbv_2 family attributes
    float lfexp;        // exp(clamp(-(Ac-2500)/50, -10, 0)): lateral-flux factor for Ac >= 2500
    float nearzero;
    float dt, inv_dt;
};

// Everything the adjoint needs from the forward evaluation of the same step.
struct Tape {
    float TTe, dT;                 // effective threshold temperature, T - TTe
    float SP1, melt0, melt2, MW1, rf0, rf2, SP3, ts0;
    float r, sw0, sw, W, infil;    // soil wetness
    float s_base, pw, fcap, fmin;  // infiltration (hourly)
    float ex0, SM2, den, ef0, ef1, ef, et1, SMd;
    float r2, capf, c1, SLZin, SM3, capillary, SMc, SLZc;  // capillary
    float SUZ1, pc, q0a, SUZ3, SLZ1, lfarg, lfval, SLZ2;
    float SPg, MWg, SMg, SUZg, SLZg;  // guard-rail inputs (hourly)
};

// ---- math primitives --------------------------------------------------------------------
// HBV_MATH = 0: IEEE division, powf/expf/logf (libdevice, ~100 instructions per powf)
// HBV_MATH = 1: SFU path (default): x^y = ex2(y * lg2(x)), division = x * rcp(y), exp via ex2.
//   The bases are strictly positive here (SM >= nearzero, FC >= 50, 1 - s >= 0.01); measured
//   parity against the reference is reported in DESIGN.md §6 (well inside 1e-5 / 1e-4).
#ifndef HBV_MATH
#define HBV_MATH 1
#endif

__device__ __forceinline__ float lg2_approx(float a) {
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r;
}
__device__ __forceinline__ float ex2_approx(float a) {
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r;
}
__device__ __forceinline__ float rcp_approx(float a) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r;
}
#if HBV_MATH == 0
__device__ __forceinline__ float pow_pos(float a, float b) { return powf(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float flog(float a) { return logf(a); }
__device__ __forceinline__ float fexp(float a) { return expf(a); }
#else
__device__ __forceinline__ float pow_pos(float a, float b) { return ex2_approx(b * lg2_approx(a)); }
__device__ __forceinline__ float fdiv(float a, float b) { return a * rcp_approx(b); }
__device__ __forceinline__ float flog(float a) { return 0.693147180559945f * lg2_approx(a); }
__device__ __forceinline__ float fexp(float a) { return ex2_approx(1.442695040888963f * a); }
#endif

// ---- the step as three stages --------------------------------------------------------------
// One HBV step is a feed-forward chain of three stages, each owning its states:
//   snow  (SNOWPACK, MELTWATER)  -> RAIN, tosoil
//   soil  (SM; reads SLZ through capillary rise in 1.1p / 2.0) -> recharge, excess, SLZa
//   resp  (SUZ, SLZ)             -> Q0, Q1, Q2, PERC
// The snow routine never reads soil or groundwater state, the soil routine reads the lower zone
// only at its very end (capillary rise), so a kernel may run snow(t+2), soil(t+1) and resp(t) in
// the same loop iteration as three independent dependency chains (hbv_pipe.cu); step_fwd /
// step_bwd below run them back to back for one t.  Arithmetic and evaluation order inside a
// stage follow the reference line by line.
struct SoilOut { float recharge, excess, ET, ef, capillary, IE; };
struct RespOut { float Q0, Q1, Q2, PERC; };

// snow routine: hbv.py:428-459 (hourly: hbv_2_hourly.py:528-566).  SP, MW updated in place
// (SP on exit is also the SWE series value).
template <int VAR, bool TAPE>
__device__ __forceinline__ void snow_fwd(float& SPio, float& MWio, const float (&p)[Traits<VAR>::NPAR],
                                         float P, float T, const LaneConst& c,
                                         float& RAIN, float& tosoil, Tape& tp) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt;
    float SP = SPio, MW = MWio;
    if constexpr (TR::HOURLY) {  // hbv_2_hourly.py:528-533
        if constexpr (TAPE) { tp.SPg = SP; tp.MWg = MW; }
        SP = fmaxf(SP, 0.f); MW = fmaxf(MW, 0.f);
    }
    float TTe = p[HBV_P_TT];
    if constexpr (TR::LAT) TTe = (c.Elev >= 2000.f) ? 4.0f : p[HBV_P_TT];  // hbv_2.py:473-475
    const bool israin = (T >= TTe);
    RAIN = israin ? P : 0.f;
    const float SNOW = (T < TTe) ? P : 0.f;
    const float dT = T - TTe;

    float SP1, melt0, melt2, melt, MW1, SP2, rf0, rf2, rf, SP3, MW2, ts0, MW3;
    if constexpr (TR::HOURLY) SP1 = SP + SNOW * dt; else SP1 = SP + SNOW;
    melt0 = p[HBV_P_CFMAX] * dT;
    melt2 = fmaxf(melt0, 0.f);
    if constexpr (TR::HOURLY) melt2 = melt2 * dt;
    melt = fminf(melt2, SP1);
    MW1 = MW + melt;
    SP2 = SP1 - melt;
    rf0 = p[HBV_P_CFR] * p[HBV_P_CFMAX] * (TTe - T);
    rf2 = fmaxf(rf0, 0.f);
    if constexpr (TR::HOURLY) rf2 = rf2 * dt;
    rf = fminf(rf2, MW1);
    SP3 = SP2 + rf;
    MW2 = MW1 - rf;
    ts0 = MW2 - p[HBV_P_CWH] * SP3;
    if constexpr (TR::HOURLY) ts0 = ts0 * inv_dt;
    tosoil = fmaxf(ts0, 0.f);
    if constexpr (TR::HOURLY) MW3 = MW2 - tosoil * dt; else MW3 = MW2 - tosoil;
    SPio = SP3; MWio = MW3;
    if constexpr (TAPE) {
        tp.TTe = TTe; tp.dT = dT;
        tp.SP1 = SP1; tp.melt0 = melt0; tp.melt2 = melt2; tp.MW1 = MW1; tp.rf0 = rf0; tp.rf2 = rf2;
        tp.SP3 = SP3; tp.ts0 = ts0;
    }
}

// soil routine + capillary rise: hbv.py:462-480, hbv_1_1p.py:482-490, hbv_2_hourly.py:568-617.
// SM updated in place; SLZ (variants with capillary rise only) enters as the lower-zone storage
// at the start of the step and leaves as SLZ after capillary rise (the response stage's input).
template <int VAR, bool BETAET, bool TAPE>
__device__ __forceinline__ void soil_fwd(float& SMio, float& SLZio, const float (&p)[Traits<VAR>::NPAR],
                                         float RAIN, float tosoil, float PET, const LaneConst& c,
                                         SoilOut& o, Tape& tp) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt, nz = c.nearzero;
    float SM = SMio, SLZ = SLZio;
    if constexpr (TR::HOURLY) {
        if constexpr (TAPE) { tp.SMg = SM; tp.SLZg = SLZ; }
        SM = fmaxf(SM, nz); SLZ = fmaxf(SLZ, nz);
    }
    const float r = fdiv(SM, p[HBV_P_FC]);
    const float sw0 = pow_pos(r, p[HBV_P_BETA]);
    const float sw = fminf(fmaxf(sw0, 0.f), 1.f);
    const float W = RAIN + tosoil;
    float infil = W, IE = 0.f, s_base = 1.f, pw = 0.f, fcap = 0.f, fmin = 0.f;
    float recharge, SM1;
    if constexpr (TR::HOURLY) {  // hbv_2_hourly.py:575-595
        const float s = fminf(fmaxf(r, 0.f), 0.99f);
        fmin = p[HBV_P_FMIN] * p[HBV_P_F0];
        s_base = 1.0f - s;
        pw = pow_pos(s_base, p[HBV_P_ALPHA]);
        fcap = fmin + (p[HBV_P_F0] - fmin) * pw;
        infil = fminf(W, fcap);
        IE = fmaxf(W - fcap, 0.f);
        recharge = infil * sw;
        SM1 = SM + (infil - recharge) * dt;
    } else {
        recharge = W * sw;
        SM1 = SM + RAIN + tosoil - recharge;
    }
    float ex0 = SM1 - p[HBV_P_FC];
    if constexpr (TR::HOURLY) ex0 = ex0 * inv_dt;
    const float excess = fmaxf(ex0, 0.f);
    float SM2;
    if constexpr (TR::HOURLY) SM2 = SM1 - excess * dt; else SM2 = SM1 - excess;
    const float den = p[HBV_P_LP] * p[HBV_P_FC];
    const float ef0 = fdiv(SM2, den);
    float ef1 = ef0;
    if constexpr (BETAET) ef1 = pow_pos(ef0, p[HBV_P_BETAET]);
    const float ef = fminf(fmaxf(ef1, 0.f), 1.f);
    float et1 = PET * ef;
    if constexpr (TR::HOURLY) et1 = et1 * dt;
    const float et2 = fminf(SM2, et1);
    float ET = et2, SMd;
    if constexpr (TR::HOURLY) { ET = et2 * inv_dt; SMd = SM2 - ET * dt; } else SMd = SM2 - ET;
    const float SM3 = fmaxf(SMd, nz);

    // Capillary rise (hbv_1_1p.py:482-490)
    float SMf = SM3, SLZa = SLZ, capillary = 0.f, r2 = 0.f, capf = 0.f, c1 = 0.f, SMc = 0.f, SLZc = 0.f;
    if constexpr (TR::CAP) {
        r2 = fdiv(SM3, p[HBV_P_FC]);
        capf = 1.0f - fminf(r2, 1.0f);
        c1 = p[HBV_P_C] * SLZ * capf;
        if constexpr (TR::HOURLY) c1 = c1 * dt;
        const float c2 = fminf(SLZ, c1);
        capillary = c2;
        if constexpr (TR::HOURLY) {
            capillary = c2 * inv_dt;
            SMc = SM3 + capillary * dt; SLZc = SLZ - capillary * dt;
        } else {
            SMc = SM3 + capillary; SLZc = SLZ - capillary;
        }
        SMf = fmaxf(SMc, nz);
        SLZa = fmaxf(SLZc, nz);
    }
    SMio = SMf;
    if constexpr (TR::CAP) SLZio = SLZa;
    o.recharge = recharge; o.excess = excess; o.ET = ET; o.ef = ef; o.capillary = capillary; o.IE = IE;
    if constexpr (TAPE) {
        tp.r = r; tp.sw0 = sw0; tp.sw = sw; tp.W = W; tp.infil = infil;
        tp.s_base = s_base; tp.pw = pw; tp.fcap = fcap; tp.fmin = fmin;
        tp.ex0 = ex0; tp.SM2 = SM2; tp.den = den; tp.ef0 = ef0; tp.ef1 = ef1; tp.ef = ef;
        tp.et1 = et1; tp.SMd = SMd;
        tp.r2 = r2; tp.capf = capf; tp.c1 = c1; tp.SLZin = SLZ; tp.SM3 = SM3;
        tp.capillary = capillary; tp.SMc = SMc; tp.SLZc = SLZc;
    }
}

// response boxes: hbv.py:483-492, hbv_2.py:545-550 (lateral flux), hbv_2_hourly.py:619-648.
// SUZ, SLZ updated in place (SLZ enters as the value after capillary rise).
template <int VAR, bool TAPE>
__device__ __forceinline__ void resp_fwd(float& SUZio, float& SLZio, const float (&p)[Traits<VAR>::NPAR],
                                         float recharge, float excess, const LaneConst& c,
                                         RespOut& o, Tape& tp) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt, nz = c.nearzero;
    float SUZ = SUZio;
    const float SLZa = SLZio;
    if constexpr (TR::HOURLY) {
        if constexpr (TAPE) tp.SUZg = SUZ;
        SUZ = fmaxf(SUZ, nz);
    }
    float SUZ1, pc, PERC, SUZ2, q0a, Q0, SUZ3, Q1, SUZ4, SLZ1;
    if constexpr (TR::HOURLY) {
        SUZ1 = SUZ + (recharge + excess) * dt;
        pc = p[HBV_P_PERC] * dt;
        PERC = fminf(SUZ1, pc) * inv_dt;
        SUZ2 = SUZ1 - PERC * dt;
    } else {
        SUZ1 = SUZ + recharge + excess;
        pc = p[HBV_P_PERC];
        PERC = fminf(SUZ1, pc);
        SUZ2 = SUZ1 - PERC;
    }
    q0a = SUZ2 - p[HBV_P_UZL];
    Q0 = p[HBV_P_K0] * fmaxf(q0a, 0.f);
    if constexpr (TR::HOURLY) SUZ3 = SUZ2 - Q0 * dt; else SUZ3 = SUZ2 - Q0;
    Q1 = p[HBV_P_K1] * SUZ3;
    if constexpr (TR::HOURLY) { SUZ4 = SUZ3 - Q1 * dt; SLZ1 = SLZa + PERC * dt; }
    else { SUZ4 = SUZ3 - Q1; SLZ1 = SLZa + PERC; }
    float SLZ2 = SLZ1, lfarg = 0.f, lfval = 0.f;
    if constexpr (TR::LAT) {  // hbv_2.py:545-550
        float LF;
        if (c.Ac < 2500.f) {
            lfarg = (c.Ac - p[HBV_P_AC]) * 0.001f;
            lfval = fminf(fmaxf(lfarg, -1.f), 1.f);
        } else {
            lfval = c.lfexp;
        }
        LF = lfval * p[HBV_P_RT];
        if constexpr (TR::HOURLY) LF = LF * dt;
        lfarg = (c.Ac < 2500.f) ? lfarg : 0.f;
        SLZ2 = fmaxf(SLZ1 + LF, 0.f);
        if constexpr (TAPE) tp.SLZ2 = SLZ1 + LF;  // pre-clamp value
    }
    const float Q2 = p[HBV_P_K2] * SLZ2;
    float SLZ3;
    if constexpr (TR::HOURLY) SLZ3 = SLZ2 - Q2 * dt; else SLZ3 = SLZ2 - Q2;
    SUZio = SUZ4; SLZio = SLZ3;
    o.Q0 = Q0; o.Q1 = Q1; o.Q2 = Q2; o.PERC = PERC;
    if constexpr (TAPE) {
        tp.SUZ1 = SUZ1; tp.pc = pc; tp.q0a = q0a; tp.SUZ3 = SUZ3; tp.SLZ1 = SLZ1;
        tp.lfarg = lfarg; tp.lfval = lfval;
        if constexpr (!TR::LAT) tp.SLZ2 = SLZ2;
    }
}

// Forward step.  S = {SNOWPACK, MELTWATER, SM, SUZ, SLZ} updated in place.
// P and PET are the forcing values as the step uses them (hourly: already / dt).
// F[] receives the NFLUX per-lane fluxes in HBV_F_* order.
template <int VAR, bool BETAET, bool TAPE>
__device__ __forceinline__ void step_fwd(float (&S)[5], const float (&p)[Traits<VAR>::NPAR],
                                         float P, float T, float PET, const LaneConst& c,
                                         float (&F)[HBV_MAX_FLUX], Tape& tp) {
    using TR = Traits<VAR>;
    float RAIN, tosoil;
    SoilOut so;
    RespOut ro;
    snow_fwd<VAR, TAPE>(S[0], S[1], p, P, T, c, RAIN, tosoil, tp);
    soil_fwd<VAR, BETAET, TAPE>(S[2], S[4], p, RAIN, tosoil, PET, c, so, tp);
    resp_fwd<VAR, TAPE>(S[3], S[4], p, so.recharge, so.excess, c, ro, tp);
    float Qsim = ro.Q0 + ro.Q1 + ro.Q2;
    if constexpr (TR::HOURLY) Qsim = Qsim + so.IE;
    F[HBV_F_QSIM] = Qsim; F[HBV_F_Q0] = ro.Q0; F[HBV_F_Q1] = ro.Q1; F[HBV_F_Q2] = ro.Q2;
    F[HBV_F_AET] = so.ET; F[HBV_F_SWE] = S[0]; F[HBV_F_RECHARGE] = so.recharge; F[HBV_F_EXCS] = so.excess;
    F[HBV_F_EVAPFACTOR] = so.ef; F[HBV_F_TOSOIL] = tosoil; F[HBV_F_PERC] = ro.PERC;
    F[HBV_F_CAPILLARY] = so.capillary;
}

// d/d{a,b} of min(a, b) with PyTorch's tie rule: returns weight on `a` (b gets 1 - w).
__device__ __forceinline__ float min_w(float a, float b) {
    return (a < b) ? 1.f : ((a > b) ? 0.f : 0.5f);
}

// QO (template flag of the adjoint stages): the caller's only cotangent is on the streamflow
// series (gF[HBV_F_QSIM]); every other gF entry is a literal zero there, but `0.f + x` is not `x`
// under IEEE rules (-0), so the compiler keeps the additions — 11 instructions per step in K2s.
template <bool QO>
__device__ __forceinline__ float cot_add(float g, float x) { return QO ? x : g + x; }

// ---- adjoint stages (reverse order: resp -> soil -> snow) -------------------------------------
// resp_bwd: in  gSUZ (= dL/dSUZ after the step), gSLZ (= dL/dSLZ after the step), gF;
//           out gSUZ (before the step), gSLZ = dL/dSLZa (the soil stage's output), and
//           gRE = d(SUZ1)-path gradient shared by recharge and excess (already "* dt").
template <int VAR, bool QO = false>
__device__ __forceinline__ void resp_bwd(float& gSUZ, float& gSLZ, const float (&gF)[HBV_MAX_FLUX],
                                         const float (&p)[Traits<VAR>::NPAR], const LaneConst& c,
                                         const Tape& tp, float (&gp)[Traits<VAR>::NPAR], float& gRE) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt, nz = c.nearzero;
    auto D = [&](float x) { return TR::HOURLY ? x * dt : x; };       // "* dt"
    auto ID = [&](float x) { return TR::HOURLY ? x * inv_dt : x; };  // "/ dt"
    const float gQ0 = cot_add<QO>(gF[HBV_F_Q0], gF[HBV_F_QSIM]);
    const float gQ1 = cot_add<QO>(gF[HBV_F_Q1], gF[HBV_F_QSIM]);
    const float gQ2 = cot_add<QO>(gF[HBV_F_Q2], gF[HBV_F_QSIM]);
    const float gSUZ4 = gSUZ;
    const float gSLZ3 = gSLZ;
    // SLZ3 = SLZ2 - Q2*dt ; Q2 = K2*SLZ2
    const float SLZ2c = TR::LAT ? fmaxf(tp.SLZ2, 0.f) : tp.SLZ2;
    const float gQ2t = gQ2 - D(gSLZ3);
    gp[HBV_P_K2] += gQ2t * SLZ2c;
    float gSLZ2 = gSLZ3 + gQ2t * p[HBV_P_K2];
    float gSLZ1 = gSLZ2;
    if constexpr (TR::LAT) {
        const float gpre = (tp.SLZ2 >= 0.f) ? gSLZ2 : 0.f;
        gSLZ1 = gpre;
        const float gLF = D(gpre);
        gp[HBV_P_RT] += gLF * tp.lfval;
        if (c.Ac < 2500.f) {
            const bool inside = (tp.lfarg >= -1.f) && (tp.lfarg <= 1.f);
            if (inside) gp[HBV_P_AC] -= gLF * p[HBV_P_RT] * (1.0f / 1000.f);
        }
    }
    // SLZ1 = SLZa + PERC*dt
    float gPERC = cot_add<QO>(gF[HBV_F_PERC], D(gSLZ1));
    // SUZ4 = SUZ3 - Q1*dt ; Q1 = K1*SUZ3
    const float gQ1t = gQ1 - D(gSUZ4);
    gp[HBV_P_K1] += gQ1t * tp.SUZ3;
    const float gSUZ3 = gSUZ4 + gQ1t * p[HBV_P_K1];
    // SUZ3 = SUZ2 - Q0*dt ; Q0 = K0*max(q0a, 0)
    const float gQ0t = gQ0 - D(gSUZ3);
    gp[HBV_P_K0] += gQ0t * fmaxf(tp.q0a, 0.f);
    const float gq0a = (tp.q0a >= 0.f) ? gQ0t * p[HBV_P_K0] : 0.f;
    const float gSUZ2 = gSUZ3 + gq0a;
    gp[HBV_P_UZL] -= gq0a;
    // SUZ2 = SUZ1 - PERC*dt ; PERC = min(SUZ1, pc)/dt ; pc = parPERC*dt
    gPERC -= D(gSUZ2);
    const float gmin = ID(gPERC);
    const float wS = min_w(tp.SUZ1, tp.pc);
    const float gSUZ1 = gSUZ2 + gmin * wS;
    gp[HBV_P_PERC] += D(gmin * (1.f - wS));
    // SUZ1 = SUZ + (recharge + excess)*dt
    gRE = D(gSUZ1);
    gSUZ = gSUZ1;
    if constexpr (TR::HOURLY) gSUZ = (tp.SUZg >= nz) ? gSUZ : 0.f;   // guard rail
    gSLZ = gSLZ1;
}

// soil_bwd: in  gSM (= dL/dSM after the step), gSLZa (from resp_bwd), gRE, gF;
//           out gSM (before the step), gSLZ (= dL/dSLZ at the start of the step: variants without
//           capillary rise pass gSLZa through), gW = dL/d(RAIN + tosoil), gPET.
template <int VAR, bool BETAET, bool QO = false>
__device__ __forceinline__ void soil_bwd(float& gSM, float& gSLZ, float gRE, const float (&gF)[HBV_MAX_FLUX],
                                         const float (&p)[Traits<VAR>::NPAR], float PET, const LaneConst& c,
                                         const Tape& tp, float (&gp)[Traits<VAR>::NPAR],
                                         float& gW, float& gPET) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt, nz = c.nearzero;
    auto D = [&](float x) { return TR::HOURLY ? x * dt : x; };
    auto ID = [&](float x) { return TR::HOURLY ? x * inv_dt : x; };
    const float gIE = gF[HBV_F_QSIM];
    const float gSMf = gSM;
    const float gSLZa = gSLZ;
    float gRech = cot_add<QO>(gF[HBV_F_RECHARGE], gRE);
    float gExc = cot_add<QO>(gF[HBV_F_EXCS], gRE);

    // Capillary
    float gSM3, gSLZ_in;
    if constexpr (TR::CAP) {
        const float ga = (tp.SLZc >= nz) ? gSLZa : 0.f;   // SLZa = max(SLZ - cap*dt, nz)
        const float gb = (tp.SMc >= nz) ? gSMf : 0.f;     // SMf  = max(SM3 + cap*dt, nz)
        gSLZ_in = ga;
        gSM3 = gb;
        const float gcap = QO ? D(gb) - D(ga) : gF[HBV_F_CAPILLARY] - D(ga) + D(gb);
        const float gc2 = ID(gcap);                       // capillary = c2/dt
        const float wL = min_w(tp.SLZin, tp.c1);          // c2 = min(SLZ, c1)
        gSLZ_in += gc2 * wL;
        const float gc0 = D(gc2 * (1.f - wL));            // c1 = c0*dt
        // c0 = (C*SLZ)*capf ; capf = 1 - min(r2, 1) ; r2 = SM3/FC
        gp[HBV_P_C] += gc0 * tp.SLZin * tp.capf;
        gSLZ_in += gc0 * p[HBV_P_C] * tp.capf;
        const float gcapf = gc0 * p[HBV_P_C] * tp.SLZin;
        const float gr2 = (tp.r2 <= 1.0f) ? -gcapf : 0.f;
        const float q = fdiv(gr2, p[HBV_P_FC]);
        gSM3 += q;
        gp[HBV_P_FC] -= q * tp.r2;
    } else {
        gSM3 = gSMf;
        gSLZ_in = gSLZa;
    }
    // SM3 = max(SMd, nz) ; SMd = SM2 - ET*dt
    const float gd = (tp.SMd >= nz) ? gSM3 : 0.f;
    float gSM2 = gd;
    const float gET = QO ? -D(gd) : gF[HBV_F_AET] - D(gd);
    // ET = et2/dt ; et2 = min(SM2, et1) ; et1 = PET*ef*dt
    const float get2 = ID(gET);
    const float wE = min_w(tp.SM2, tp.et1);
    gSM2 += get2 * wE;
    const float get0 = D(get2 * (1.f - wE));
    const float gef = QO ? get0 * PET : gF[HBV_F_EVAPFACTOR] + get0 * PET;
    gPET = get0 * tp.ef;                                   // d/dPET (as the step uses it)
    const float gef1 = (tp.ef1 >= 0.f && tp.ef1 <= 1.f) ? gef : 0.f;
    float gef0 = gef1;
    if constexpr (BETAET) {
        gef0 = fdiv(gef1 * p[HBV_P_BETAET] * tp.ef1, tp.ef0);
        gp[HBV_P_BETAET] += gef1 * tp.ef1 * flog(tp.ef0);
    }
    // ef0 = SM2/den ; den = LP*FC
    const float qe = fdiv(gef0, tp.den);
    gSM2 += qe;
    const float gden = -qe * tp.ef0;
    gp[HBV_P_LP] += gden * p[HBV_P_FC];
    gp[HBV_P_FC] += gden * p[HBV_P_LP];
    // SM2 = SM1 - excess*dt ; excess = max(ex0, 0) ; ex0 = (SM1 - FC)/dt
    float gSM1 = gSM2;
    const float gexc = gExc - D(gSM2);
    const float gex0 = (tp.ex0 >= 0.f) ? ID(gexc) : 0.f;
    gSM1 += gex0;
    gp[HBV_P_FC] -= gex0;
    // SM1 = SM + (infil - recharge)*dt
    float gSM_in = gSM1;
    float ginfil = D(gSM1);
    gRech -= D(gSM1);
    // recharge = infil*sw
    ginfil += gRech * tp.sw;
    const float gsw = gRech * tp.infil;
    float gr = 0.f;
    if constexpr (TR::HOURLY) {
        // infil = min(W, fcap) ; IE = max(W - fcap, 0)
        const float wW = min_w(tp.W, tp.fcap);
        const float ge = (tp.W - tp.fcap >= 0.f) ? gIE : 0.f;
        gW = ginfil * wW + ge;
        const float gfcap = ginfil * (1.f - wW) - ge;
        // fcap = fmin + (F0 - fmin)*pw ; fmin = FMIN*F0 ; pw = s_base^ALPHA
        const float gfmin = gfcap * (1.f - tp.pw);
        const float gpw = gfcap * (p[HBV_P_F0] - tp.fmin);
        gp[HBV_P_F0] += gfcap * tp.pw + gfmin * p[HBV_P_FMIN];
        gp[HBV_P_FMIN] += gfmin * p[HBV_P_F0];
        gp[HBV_P_ALPHA] += gpw * tp.pw * flog(tp.s_base);
        const float gbase = fdiv(gpw * p[HBV_P_ALPHA] * tp.pw, tp.s_base);
        gr = (tp.r >= 0.f && tp.r <= 0.99f) ? -gbase : 0.f;
    } else {
        gW = ginfil;
    }
    // sw = clamp(sw0, 0, 1) ; sw0 = r^BETA ; r = SM/FC
    const float gsw0 = (tp.sw0 >= 0.f && tp.sw0 <= 1.f) ? gsw : 0.f;
    gr += fdiv(gsw0 * p[HBV_P_BETA] * tp.sw0, tp.r);
    gp[HBV_P_BETA] += gsw0 * tp.sw0 * flog(tp.r);
    const float qr = fdiv(gr, p[HBV_P_FC]);
    gSM_in += qr;
    gp[HBV_P_FC] -= qr * tp.r;
    gSM = gSM_in;
    gSLZ = gSLZ_in;
    if constexpr (TR::HOURLY) {  // guard rails: clamp(x, min=m) passes where x >= m
        gSM = (tp.SMg >= nz) ? gSM : 0.f;
        gSLZ = (tp.SLZg >= nz) ? gSLZ : 0.f;
    }
}

// snow_bwd: in  gSP (= dL/dSNOWPACK after the step, the SWE series' cotangent NOT yet added),
//           gMW (after the step), gW (from soil_bwd), gF; out gSP, gMW before the step,
//           gP = dL/dP, gT = dL/dT.
template <int VAR, bool QO = false>
__device__ __forceinline__ void snow_bwd(float& gSP, float& gMW, float gW, const float (&gF)[HBV_MAX_FLUX],
                                         const float (&p)[Traits<VAR>::NPAR], const LaneConst& c,
                                         const Tape& tp, float (&gp)[Traits<VAR>::NPAR],
                                         float& gP, float& gT) {
    using TR = Traits<VAR>;
    const float dt = c.dt, inv_dt = c.inv_dt;
    auto D = [&](float x) { return TR::HOURLY ? x * dt : x; };
    auto ID = [&](float x) { return TR::HOURLY ? x * inv_dt : x; };
    float gSP3 = cot_add<QO>(gF[HBV_F_SWE], gSP);
    const float gMW3 = gMW;
    // W = RAIN + tosoil ; MW3 = MW2 - tosoil*dt ; tosoil = max(ts0, 0) ; ts0 = (MW2 - CWH*SP3)/dt
    const float gtosoil = QO ? gW - D(gMW3) : gF[HBV_F_TOSOIL] + gW - D(gMW3);
    const float gx = (tp.ts0 >= 0.f) ? ID(gtosoil) : 0.f;
    const float gMW2 = gMW3 + gx;
    gp[HBV_P_CWH] -= gx * tp.SP3;
    gSP3 -= gx * p[HBV_P_CWH];
    // SP3 = SP2 + rf ; MW2 = MW1 - rf ; rf = min(rf2, MW1) ; rf2 = max(rf0, 0)*dt
    const float gSP2 = gSP3;
    const float grf = gSP3 - gMW2;
    const float wR = min_w(tp.rf2, tp.MW1);
    const float gMW1 = gMW2 + grf * (1.f - wR);
    const float grf0 = (tp.rf0 >= 0.f) ? D(grf * wR) : 0.f;
    // rf0 = CFR*CFMAX*(TTe - T)
    float gTTe = grf0 * p[HBV_P_CFR] * p[HBV_P_CFMAX];
    gp[HBV_P_CFR] += grf0 * p[HBV_P_CFMAX] * (-tp.dT);
    gp[HBV_P_CFMAX] += grf0 * p[HBV_P_CFR] * (-tp.dT);
    // MW1 = MW + melt ; SP2 = SP1 - melt ; melt = min(melt2, SP1) ; melt2 = max(melt0, 0)*dt
    const float gMW_in = gMW1;
    const float gmelt = gMW1 - gSP2;
    const float wM = min_w(tp.melt2, tp.SP1);
    const float gSP1 = gSP2 + gmelt * (1.f - wM);
    const float gmelt0 = (tp.melt0 >= 0.f) ? D(gmelt * wM) : 0.f;
    gp[HBV_P_CFMAX] += gmelt0 * tp.dT;
    gTTe -= gmelt0 * p[HBV_P_CFMAX];
    if constexpr (TR::LAT) { if (c.Elev < 2000.f) gp[HBV_P_TT] += gTTe; }
    else gp[HBV_P_TT] += gTTe;
    // forcings: T enters as T - TTe only (the rain/snow masks carry no gradient);
    // P = RAIN (T >= TTe, into W) or SNOW (T < TTe, into SP1 = SP + SNOW*dt)
    gT = -gTTe;
    gP = (tp.dT >= 0.f) ? gW : D(gSP1);
    gSP = gSP1; gMW = gMW_in;
    if constexpr (TR::HOURLY) {
        gSP = (tp.SPg >= 0.f) ? gSP : 0.f;
        gMW = (tp.MWg >= 0.f) ? gMW : 0.f;
    }
}

// Adjoint step.  On entry gS = dL/d(state after the step); gF = dL/d(per-lane fluxes of the
// step).  On exit gS = dL/d(state before the step), gp[i] += dL/d(parameter i at this step) and
// gX = dL/d(P, T, PET) of this step as the step uses them (hourly: P, PET already / dt).
template <int VAR, bool BETAET, bool QO = false>
__device__ __forceinline__ void step_bwd(float (&gS)[5], const float (&gF)[HBV_MAX_FLUX],
                                         const float (&p)[Traits<VAR>::NPAR], float PET,
                                         const LaneConst& c, const Tape& tp,
                                         float (&gp)[Traits<VAR>::NPAR], float (&gX)[3]) {
    float gRE, gW;
    resp_bwd<VAR, QO>(gS[3], gS[4], gF, p, c, tp, gp, gRE);
    soil_bwd<VAR, BETAET, QO>(gS[2], gS[4], gRE, gF, p, PET, c, tp, gp, gW, gX[2]);
    snow_bwd<VAR, QO>(gS[0], gS[1], gW, gF, p, c, tp, gp, gX[0], gX[1]);
}

__device__ __forceinline__ void init_lane_const(LaneConst& lc, float Ac, float Elev) {
    lc.Ac = Ac; lc.Elev = Elev;
    lc.lfexp = expf(fminf(fmaxf(-(Ac - 2500.f) / 50.f, -10.f), 0.f));   // hbv_2.py:547-549
}

// sigmoid + affine descale (hbv.py:201, core/calc/utils.py:24)
// sigmoid (hbv.py:201).  Kept at full float accuracy in both math modes: the threshold
// temperature enters as T - parTT, a cancellation that amplifies parameter error ~10x.
__device__ __forceinline__ float sigmoidf_(float x) {
    const float e = expf(-fminf(fmaxf(x, -80.f), 80.f));
    const float dn = 1.0f + e;
    float r = rcp_approx(dn);
    r = fmaf(r, fmaf(-dn, r, 1.0f), r);   // one Newton step: ~0.5 ulp
    return r;
}

// SFU sigmoid, 4 instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP), ~3 ulp: used for every parameter
// except parTT (see above).  With all 14 parameters of hbv_1_1p time-varying the full-accuracy
// form costs ~250 of ~470 warp instructions per step; this one ~60.
__device__ __forceinline__ float sigmoid_sfu(float x) {
#if HBV_MATH == 0
    return sigmoidf_(x);
#else
    return rcp_approx(1.0f + ex2_approx(-1.442695040888963f * x));
#endif
}

}  // namespace hbv
