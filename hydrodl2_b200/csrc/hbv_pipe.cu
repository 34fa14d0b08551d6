// hbv_pipe.cu — K1p / K2p: the recurrence and its every-state adjoint as a software pipeline of
// the three HBV stages, for the latency-bound regime (a few warps per SM scheduler: 531 basins,
// or 2,500 hourly units x 17,520 steps per GPU).
//
// Round 1's ncu captures (profiles/r01_ncu_c2_lean.md) showed the one-warp kernels at 0.33-0.47
// issued instructions per cycle with `wait` / `short_scoreboard` as the stall reasons: one step is
// ONE dependent chain snow -> soil -> response (hbv.py:428-492), ~400 cycles for ~150
// instructions, and nothing else is resident on the scheduler to fill the gaps.  But the chain is
// longer than it needs to be: the snow routine (hbv.py:428-459) never reads soil or groundwater
// state, the soil routine (hbv.py:462-480) reads the lower zone only through capillary rise at
// its very end (hbv_1_1p.py:482-490), and the response boxes (hbv.py:483-492) only consume
// recharge + excess.  So loop iteration k runs
//        forward:   resp(k-2)      soil(k-1)      snow(k)
//        adjoint:   fwd-recompute(i)   soil_bwd(i+1)   resp_bwd(i)   snow_bwd(i+2)     (i descending)
// as independent instruction streams inside one straight-line block: the per-step critical path
// is the longest STAGE (soil: two pow = four MUFU + ~17 dependent FP32 ops), not the sum, and the
// scheduler's issue slots fill up from ~35 % towards the instruction count.  The few values that
// cross a stage boundary (RAIN, tosoil, PET, recharge, excess; the tape of a step in the adjoint)
// are carried in registers; loops are unrolled by the carry distance so the carries are register
// renames, not moves.  Same hbv_step.cuh arithmetic as every other kernel family.
//
// Layout / eligibility as hbv_lean.cu (standard 3-wide x_phy, nmul 16, compile-time dynamic set,
// loss on the streamflow series, K = 1); one-warp CTAs (2 basins x 16 components), inputs staged
// through a cp.async ring.  Reference spans replaced: models/hbv/hbv.py:423-511,
// hbv_1_1p.py:422-524, hbv_2.py:464-585, hbv_2_hourly.py:527-683 and autograd over them.
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "hbv_common.cuh"

namespace hbv {

constexpr int PNM = 16;      // components per basin
constexpr int PBPB = 2;      // basins per (one-warp) CTA
constexpr int PTC = 4;       // forward: time steps per output chunk

// pipeline stage that consumes parameter i: 0 = snow, 1 = soil, 2 = response
__host__ __device__ constexpr int par_stage(int i) {
    return (i == HBV_P_TT || i == HBV_P_CFMAX || i == HBV_P_CFR || i == HBV_P_CWH) ? 0
         : (i == HBV_P_FC || i == HBV_P_BETA || i == HBV_P_LP || i == HBV_P_BETAET || i == HBV_P_C ||
            i == HBV_P_F0 || i == HBV_P_FMIN || i == HBV_P_ALPHA) ? 1 : 2;
}

template <int NPAR, int DM, int LAYOUT>
__host__ __device__ constexpr int pipe_col(int i) {
    return (LAYOUT == 0 ? i : DynSet<NPAR, DM>::slot(i)) * PNM;
}

template <bool SIG>
__device__ __forceinline__ float pipe_descale(int i, float raw, float span, float lo) {
    if constexpr (SIG) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        return fmaf(s, span, lo);
    } else {
        return fmaf(raw, span, lo);
    }
}

template <bool SIG>
__device__ __forceinline__ void pipe_descale_both(int i, float raw, float span, float lo, float& v, float& dv) {
    if constexpr (SIG) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        v = fmaf(s, span, lo);
        dv = span * s * (1.0f - s);
    } else {
        v = fmaf(raw, span, lo);
        dv = span;
    }
}

template <int U> using IC = std::integral_constant<int, U>;

// ================================================================================================
// K1p: forward.  CK: store the state before every step (K = 1); WF: write the flux planes.
// ================================================================================================
template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, bool CK, bool WF, int RD>
__global__ void __launch_bounds__(PBPB * PNM, 1)
hbv_fwd_pipe_kernel(const KDesc d, const FwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    constexpr int NDR = DS::NDYN;
    constexpr int ND = NDR > 0 ? NDR : 1;
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b_raw = blockIdx.x * PBPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * PNM + j;
    const int64_t nlane = (int64_t)d.B * PNM;
    const int T = d.T;

    LaneConst lc;
    lc.nearzero = d.nearzero; lc.dt = d.dt; lc.inv_dt = d.inv_dt;
    lc.Ac = 0.f; lc.Elev = 0.f; lc.lfexp = 0.f;
    if constexpr (TR::LAT) init_lane_const(lc, __ldg(io.attrs + b), __ldg(io.attrs + d.B + b));

    float p[NPAR];
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, nullptr, nullptr);
    float dspan[ND], dlo[ND];
#pragma unroll
    for (int i = 0; i < NPAR; ++i)
        if (DS::is_dyn(i, 0)) { dspan[DS::slot(i)] = d.span[i]; dlo[DS::slot(i)] = d.lo[i]; }

    float S[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) S[s] = __ldg(io.state_in + s * nlane + lane);

    // ---- output tile: two chunks of PTC time slots (stage s of step t writes slot t mod 8 while
    // the previous chunk is being reduced), 12 floats per lane and slot:
    //   [0..3] Qsim Q0 Q1 Q2 (resp) | [4..7] AET recharge excess evapfactor (soil) |
    //   [8] SWE [9] tosoil (snow) | [10] PERC (resp) | [11] capillary (soil)
    constexpr int bstride = PNM * NFP + 12;
    constexpr int tstride_s = PBPB * bstride;
    constexpr int TILE = WF ? 2 * PTC * tstride_s : 0;
    float* const my_slot = smem + bl * bstride + j * NFP;
    constexpr float inv_nmul = 1.0f / (float)PNM;
    // reduce item of this lane: (time slot, basin, quad); 24 items per chunk
    const int r_q = tid % 3;
    const int r_r = tid / 3;
    const int r_bl = r_r % PBPB;
    const int r_tc = r_r / PBPB;
    const int r_bb = blockIdx.x * PBPB + r_bl;
    const bool r_ok = (tid < PTC * PBPB * 3) && (r_bb < d.B);
    const int r_off = r_tc * tstride_s + r_bl * bstride + r_q * 4;
    float* pl[4] = {nullptr, nullptr, nullptr, nullptr};
    if constexpr (WF) {
        // tile quad -> flux planes
        const int f0 = r_q == 0 ? HBV_F_QSIM : (r_q == 1 ? HBV_F_AET : HBV_F_SWE);
        const int f1 = r_q == 0 ? HBV_F_Q0 : (r_q == 1 ? HBV_F_RECHARGE : HBV_F_TOSOIL);
        const int f2 = r_q == 0 ? HBV_F_Q1 : (r_q == 1 ? HBV_F_EXCS : HBV_F_PERC);
        const int f3 = r_q == 0 ? HBV_F_Q2 : (r_q == 1 ? HBV_F_EVAPFACTOR : HBV_F_CAPILLARY);
        const int64_t o = (int64_t)r_tc * d.B + (r_ok ? r_bb : 0);
        pl[0] = io.flux[f0] + o; pl[1] = io.flux[f1] + o; pl[2] = io.flux[f2] + o;
        pl[3] = (f3 < TR::NFLUX) ? io.flux[f3] + o : nullptr;
    }
    // Branch-free on purpose: every lane runs the loads and adds (lanes without an item re-read
    // item 0), only the four stores are predicated — so ptxas can interleave this block with the
    // stage arithmetic of the iteration it sits in instead of serialising a divergent region.
    const int r_off_safe = (tid < PTC * PBPB * 3) ? r_off : 0;
    auto reduce_chunk = [&](const float* base, int64_t t0, int tcn) {
        // four partial sums of four components each, then a pair tree: dependent depth 5 instead
        // of a 16-long chain, 16 accumulator registers instead of 64
        const float* src = base + r_off_safe;
        float4 v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) v[jj] = *reinterpret_cast<const float4*>(src + jj * NFP);
#pragma unroll
        for (int w = 1; w < 4; ++w) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 x = *reinterpret_cast<const float4*>(src + (w * 4 + jj) * NFP);
                v[jj].x += x.x; v[jj].y += x.y; v[jj].z += x.z; v[jj].w += x.w;
            }
        }
        v[0].x = (v[0].x + v[1].x) + (v[2].x + v[3].x);
        v[0].y = (v[0].y + v[1].y) + (v[2].y + v[3].y);
        v[0].z = (v[0].z + v[1].z) + (v[2].z + v[3].z);
        v[0].w = (v[0].w + v[1].w) + (v[2].w + v[3].w);
        const bool w_ok = r_ok && r_tc < tcn;
        const int64_t o = t0 * d.B;
        if (w_ok) {
            pl[0][o] = v[0].x * inv_nmul;
            pl[1][o] = v[0].y * inv_nmul;
            pl[2][o] = v[0].z * inv_nmul;
        }
        if (w_ok && pl[3] != nullptr) pl[3][o] = v[0].w * inv_nmul;
    };

    // ---- input ring (as hbv_lean.cu): two cp.async instructions stage a warp's inputs of a step
    //   A  4 B x 6 lanes   P, T, PET of the two basins               -> slot[4 bl + k]
    //   B  8 B x 16*ND     the 64 B run of every dynamic parameter   -> slot[8 + 32 k + lane]
    constexpr int PARB = 8, SLOT = 8 + 32 * NDR;
    constexpr int NB = (16 * NDR + 31) / 32;
    // Ring position of step k is k mod RD; RD = 3 trips of PTC steps, so inside trip m it is
    // rb + U * SLOT with rb = (m mod 3) * PTC * SLOT: one base update per trip instead of a
    // compare-and-wrap per pointer and step.
    static_assert(RD == 3 * PTC, "ring depth is three output chunks");
    float* const ring0 = smem + TILE;
    int rb = 0, rb_prev = 2 * PTC * SLOT;
    const int b0w = blockIdx.x * PBPB;
    const int64_t sf = (int64_t)d.B * 3, sd = (int64_t)d.B * d.dyn_ncol;
    const int kA = tid & 15, bbA = tid >> 4;
    const bool actA = kA < 3;
    // sources as (row-0 pointer, byte stride per step): one IMAD.WIDE.U32 per copy (hbv_lean.cu)
    const char* const baseA = reinterpret_cast<const char*>(io.forcing + (int64_t)min(b0w + bbA, d.B - 1) * 3 + kA);
    const uint32_t strA = (uint32_t)d.B * 12u, strB = (uint32_t)d.B * (uint32_t)d.dyn_ncol * 4u;
    const int dstA = 4 * bbA + kA;
    const char* baseB[NB > 0 ? NB : 1];
    int dstB[NB > 0 ? NB : 1];
    bool actB[NB > 0 ? NB : 1];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        const int gq = o * 32 + tid;
        actB[o] = gq < 16 * NDR;
        const int r = actB[o] ? (gq >> 3) : 0, q = gq & 7;
        const int bb = r / (NDR > 0 ? NDR : 1), k = r - bb * NDR;
        int col = 0;
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0) && DS::slot(i) == k) col = pipe_col<NPAR, DM, LAYOUT>(i);
        baseB[o] = reinterpret_cast<const char*>(io.dyn + (int64_t)min(b0w + bb, d.B - 1) * d.dyn_ncol + col + 2 * q);
        dstB[o] = PARB + k * 32 + bb * 16 + 2 * q;
    }
    int t_issue = 0;
    auto issue = [&](float* wp) {    // stage the next time step (the last row again past the end)
        const uint64_t ts = (uint32_t)t_issue;
        if (actA) cp_async4(wp + dstA, reinterpret_cast<const float*>(baseA + ts * strA));
#pragma unroll
        for (int o = 0; o < NB; ++o)
            if (actB[o]) cp_async8(wp + dstB[o], reinterpret_cast<const float*>(baseB[o] + ts * strB));
        cp_async_commit();
        if (t_issue + 1 < T) ++t_issue;
    };
#pragma unroll 1
    for (int q = 0; q < RD - 1; ++q) issue(ring0 + q * SLOT);      // steps 0 .. RD-2

    // ---- values that cross a stage boundary (index = age in iterations) ------------------------
    float rn1 = 0.f, ts1 = 0.f, pet1 = 0.f;          // snow(k-1) -> soil(k-1): RAIN, tosoil; PET of k-1
    float rech1 = 0.f, exc1 = 0.f, ie1 = 0.f;        // soil(k-2) -> resp(k-2)
    float d1[ND], d2[ND];                            // dynamic parameters of steps k-1, k-2
#pragma unroll
    for (int s = 0; s < ND; ++s) { d1[s] = 0.f; d2[s] = 0.f; }

    // stored-state planes: (t, s) at ckpt + (5 t + s) nlane
    float* pkA = CK ? io.ckpt + lane : nullptr;                      // (k, 0)      SNOWPACK, MELTWATER
    float* pkB = CK ? io.ckpt + lane - 3 * nlane : nullptr;          // (k - 1, 2)  SM; + 2 nlane: SLZ
    float* pkC = CK ? io.ckpt + lane - 7 * nlane : nullptr;          // (k - 2, 3)  SUZ
    const int64_t stride5 = 5 * nlane;

    Tape tp;
    int par = 0;      // tile chunk parity of the current trip
    // one loop iteration k = 4 m + U:  resp(k-2), soil(k-1), snow(k)
    auto body = [&](auto Uc, const bool sn, const bool so_, const bool rs, const bool red, const int m) {
        constexpr int U = decltype(Uc)::value;
        [[maybe_unused]] float* const cur = my_slot + par * (PTC * tstride_s);
        [[maybe_unused]] float* const prv = my_slot + (par ^ 1) * (PTC * tstride_s);
        float d0[ND];
        float P0 = 0.f, T0 = 0.f, pet0 = 0.f;
        if (sn) {
            cp_async_wait<RD - 2>();
            __syncwarp();
            // step k + RD - 1 goes where step k - 1 was
            issue(ring0 + (U >= 1 ? rb + (U - 1) * SLOT : rb_prev + (PTC - 1) * SLOT));
            const float* rp = ring0 + rb + U * SLOT;
            const float4 f = *reinterpret_cast<const float4*>(rp + 4 * bl);
            P0 = f.x; T0 = f.y; pet0 = f.z;
            if constexpr (TR::HOURLY) { P0 = P0 * d.inv_dt; pet0 = pet0 * d.inv_dt; }
#pragma unroll
            for (int i = 0; i < NPAR; ++i)
                if (DS::is_dyn(i, 0))
                    d0[DS::slot(i)] = pipe_descale<SIG>(i, rp[PARB + DS::slot(i) * 32 + tid], dspan[DS::slot(i)], dlo[DS::slot(i)]);
        } else {
#pragma unroll
            for (int s = 0; s < ND; ++s) d0[s] = d1[s];
        }
        if constexpr (WF) {
            if (U == 2 && red) {      // chunk m - 1 is complete: resp(4 m - 1) ran one iteration ago
                __syncwarp();
                reduce_chunk(smem + (par ^ 1) * (PTC * tstride_s), (int64_t)(m - 1) * PTC, min(PTC, T - (m - 1) * PTC));
            }
        }
        // every parameter is read by exactly one stage: give it that stage's time step
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0))
                p[i] = par_stage(i) == 0 ? d0[DS::slot(i)] : (par_stage(i) == 1 ? d1[DS::slot(i)] : d2[DS::slot(i)]);
        // ---- response boxes of step k - 2
        if (rs) {
            if constexpr (CK) { if (valid) *pkC = S[3]; }
            RespOut ro;
            resp_fwd<VAR, false>(S[3], S[4], p, rech1, exc1, lc, ro, tp);
            if constexpr (WF) {
                float Qsim = ro.Q0 + ro.Q1 + ro.Q2;
                if constexpr (TR::HOURLY) Qsim = Qsim + ie1;
                float* o = (U >= 2) ? cur + (U - 2) * tstride_s : prv + (U + 2) * tstride_s;
                *reinterpret_cast<float4*>(o) = make_float4(Qsim, ro.Q0, ro.Q1, ro.Q2);
                o[10] = ro.PERC;
            }
        }
        // ---- soil routine of step k - 1 (reads the lower zone resp(k-2) just produced)
        if (so_) {
            if constexpr (CK) { if (valid) { pkB[0] = S[2]; pkB[2 * nlane] = S[4]; } }
            SoilOut so;
            soil_fwd<VAR, BETAET, false>(S[2], S[4], p, rn1, ts1, pet1, lc, so, tp);
            if constexpr (WF) {
                float* o = (U >= 1) ? cur + (U - 1) * tstride_s : prv + 3 * tstride_s;
                *reinterpret_cast<float4*>(o + 4) = make_float4(so.ET, so.recharge, so.excess, so.ef);
                o[11] = so.capillary;
            }
            rech1 = so.recharge; exc1 = so.excess; ie1 = so.IE;
        }
        // ---- snow routine of step k
        if (sn) {
            if constexpr (CK) { if (valid) { pkA[0] = S[0]; pkA[nlane] = S[1]; } }
            float rn0, ts0;
            snow_fwd<VAR, false>(S[0], S[1], p, P0, T0, lc, rn0, ts0, tp);
            if constexpr (WF) {
                float* o = cur + U * tstride_s;
                *reinterpret_cast<float2*>(o + 8) = make_float2(S[0], ts0);
            }
            rn1 = rn0; ts1 = ts0; pet1 = pet0;
        }
#pragma unroll
        for (int s = 0; s < ND; ++s) { d2[s] = d1[s]; d1[s] = d0[s]; }
        if constexpr (CK) { pkA += stride5; pkB += stride5; pkC += stride5; }
    };

    auto next_trip = [&]() {
        par ^= 1;
        rb_prev = rb;
        rb = (rb == 2 * PTC * SLOT) ? 0 : rb + PTC * SLOT;
    };
    const int m_last = (T + 1) / PTC;            // trip holding iteration k = T + 1
    const int m_fast_end = T / PTC;              // trips 1 .. m_fast_end - 1: every stage active
    auto slow_trip = [&](int m) {
        const int k0 = m * PTC;
        body(IC<0>{}, k0 + 0 < T, k0 + 0 >= 1 && k0 + 0 <= T, k0 + 0 >= 2 && k0 + 0 <= T + 1, false, m);
        body(IC<1>{}, k0 + 1 < T, k0 + 1 >= 1 && k0 + 1 <= T, k0 + 1 >= 2 && k0 + 1 <= T + 1, false, m);
        body(IC<2>{}, k0 + 2 < T, k0 + 2 <= T, k0 + 2 <= T + 1, m >= 1, m);
        body(IC<3>{}, k0 + 3 < T, k0 + 3 <= T, k0 + 3 <= T + 1, false, m);
        next_trip();
    };
    int m = 0;
    slow_trip(m++);
#pragma unroll 1
    for (; m < m_fast_end; ++m) {
        body(IC<0>{}, true, true, true, false, m);
        body(IC<1>{}, true, true, true, false, m);
        body(IC<2>{}, true, true, true, true, m);
        body(IC<3>{}, true, true, true, false, m);
        next_trip();
    }
#pragma unroll 1
    for (; m <= m_last; ++m) slow_trip(m);
    cp_async_wait<0>();
    if constexpr (WF) {
        // chunks the loop has not reduced yet: m_last - 1 was reduced in trip m_last (U = 2)
        __syncwarp();
        if (m_last * PTC < T) {
            // `par` was flipped after trip m_last: its chunk sits in the other half
            reduce_chunk(smem + (par ^ 1) * (PTC * tstride_s), (int64_t)m_last * PTC, T - m_last * PTC);
        }
    }
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

// ================================================================================================
// K2p: adjoint, every state stored (K = 1), upstream gradient on the streamflow series.
//
// With every state stored, the forward re-evaluation of a step has no loop-carried input — each
// stage reads its states from the stored planes — but it is still a 300-cycle chain
// snow -> soil -> resp in front of the adjoint that needs its intermediates.  So the
// re-evaluation is stage-pipelined over the (descending) sweep, and the adjoint of a step runs
// un-skewed behind it.  Iteration i:
//     snow_fwd(i-2)   soil_fwd(i-1)   resp_fwd(i) -> resp_bwd(i) -> soil_bwd(i) -> snow_bwd(i)
// Two chains of about equal length (soil_fwd: two pow; resp_fwd + the three adjoint stages) issue
// side by side.  The tape of step t is filled over three iterations (snow fields at t+2, soil
// fields at t+1, resp fields at t) in one of three rotating register buffers; the loop is
// unrolled by three so the rotation is a renaming.  A ring slot (the inputs + stored states of
// one step) is likewise read over three iterations.
// ZF: zero the CTA's gradient rows before writing them (no memset of the dense plane).
// ================================================================================================
// (registers are left unconstrained: capping the hourly variant at 224 so that BASELINE config 4's
// 1,250 one-warp CTAs fit the SMs in one wave made ptxas spill ~90 values per iteration, and the
// kernel ran 16.7 ms against K2s' 12.3 ms — measured; pipe_fits_one_wave() sends such grids to K2s)
template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, bool ZF, int RD>
__global__ void __launch_bounds__(PBPB * PNM, 1)
hbv_bwd_pipe_kernel(const KDesc d, const BwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    constexpr int NDR = DS::NDYN;
    constexpr int ND = NDR > 0 ? NDR : 1;
    static_assert(RD == 12, "ring depth is four trips of three iterations");
    extern __shared__ __align__(16) float ringmem[];

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b_raw = blockIdx.x * PBPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * PNM + j;
    const int64_t nlane = (int64_t)d.B * PNM;
    const int T = d.T;

    LaneConst lc;
    lc.nearzero = d.nearzero; lc.dt = d.dt; lc.inv_dt = d.inv_dt;
    lc.Ac = 0.f; lc.Elev = 0.f; lc.lfexp = 0.f;
    if constexpr (TR::LAT) init_lane_const(lc, __ldg(io.attrs + b), __ldg(io.attrs + d.B + b));

    float p[NPAR], gacc[NPAR];
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < NPAR; ++i) gacc[i] = 0.f;
    float dspan[ND], dlo[ND];
#pragma unroll
    for (int i = 0; i < NPAR; ++i)
        if (DS::is_dyn(i, 0)) { dspan[DS::slot(i)] = d.span[i]; dlo[DS::slot(i)] = d.lo[i]; }

    float gS[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) gS[s] = io.gstate_out ? __ldg(io.gstate_out + s * nlane + lane) : 0.f;

    const int64_t row_last = (int64_t)(T - 1) * d.B + b;
    const int64_t sf = (int64_t)d.B * 3, sd = (int64_t)d.B * d.dyn_ncol;
    float* pg = io.gdyn + row_last * d.dyn_ncol + j;       // gradient row of step i
    constexpr float inv_nmul = 1.0f / (float)PNM;

    // ---- ring (as hbv_lean.cu's adjoint): four cp.async instructions stage a step
    //   A  4 B x 8 lanes   P, T, PET and dL/dQ of the two basins         -> slot[0 .. 7]
    //   B  8 B x 16*ND     the 64 B run of every dynamic parameter        -> slot[8 + 32 k + lane]
    //   C 16 B x 32 lanes  stored states 0..3 (128 B per state and warp)  -> slot[ST + 32 s + lane]
    //   D 16 B x 8 lanes   stored state 4
    // Steps are staged in sweep order (T-1 first): step t sits at position (T-1-t) mod RD.  With
    // n = T+1-i counting iterations, iteration i reads positions n (snow inputs of step i-2),
    // n-1 (soil inputs of step i-1), n-2 (resp inputs of step i) and refills n-3 (step i+1 is
    // done) — inside trip m (n = 3 m + U) these are rb / rb_prev + constants, rb = (m mod 4) * 3 SLOT.
    constexpr int PARB = 8, STB = 8 + 32 * NDR, SLOT = 8 + 32 * (NDR + 5);
    constexpr int NB = (16 * NDR + 31) / 32;
    int rb = 0, rb_prev = 9 * SLOT;
    const int b0w = blockIdx.x * PBPB;
    const int nbw = min(PBPB, d.B - b0w);
    const int kA = tid & 15, bbA = tid >> 4;
    const bool actA = kA < 4;
    const int64_t rowA = (int64_t)(T - 1) * d.B + min(b0w + bbA, d.B - 1);
    const float* srcA = (kA < 3) ? io.forcing + rowA * 3 + kA : io.gflux[HBV_F_QSIM] + rowA;
    const int64_t strA = (kA < 3) ? sf : (int64_t)d.B;
    const int dstA = 4 * bbA + kA;
    const float* srcB[NB > 0 ? NB : 1];
    int dstB[NB > 0 ? NB : 1];
    bool actB[NB > 0 ? NB : 1];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        const int gq = o * 32 + tid;
        actB[o] = gq < 16 * NDR;
        const int r = actB[o] ? (gq >> 3) : 0, q = gq & 7;
        const int bb = r / (NDR > 0 ? NDR : 1), k = r - bb * NDR;
        int col = 0;
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0) && DS::slot(i) == k) col = pipe_col<NPAR, DM, LAYOUT>(i);
        srcB[o] = io.dyn + ((int64_t)(T - 1) * d.B + min(b0w + bb, d.B - 1)) * d.dyn_ncol + col + 2 * q;
        dstB[o] = PARB + k * 32 + bb * 16 + 2 * q;
    }
    const int sC = tid >> 3, qC = tid & 7;
    const bool actC = 4 * qC < 16 * nbw;
    const float* srcC = io.ckpt + ((int64_t)(T - 1) * 5 + sC) * nlane + (int64_t)b0w * PNM + 4 * qC;
    const float* srcD = io.ckpt + ((int64_t)(T - 1) * 5 + 4) * nlane + (int64_t)b0w * PNM + 4 * qC;
    const bool actD = actC && tid < 8;
    const int dstC = STB + sC * 32 + 4 * qC, dstD = STB + 4 * 32 + 4 * qC;
    const int64_t strC = 5 * nlane;
    int t_stage = T - 1;
    auto issue = [&](float* wp) {    // stage the inputs of the next step of the sweep (if any)
        if (t_stage >= 0) {
            if (actA) cp_async4(wp + dstA, srcA);
#pragma unroll
            for (int o = 0; o < NB; ++o)
                if (actB[o]) cp_async8(wp + dstB[o], srcB[o]);
            if (actC) cp_async16(wp + dstC, srcC);
            if (actD) cp_async16(wp + dstD, srcD);
            srcA -= strA; srcC -= strC; srcD -= strC;
#pragma unroll
            for (int o = 0; o < NB; ++o) srcB[o] -= sd;
        }
        --t_stage;
        cp_async_commit();
    };
#pragma unroll 1
    for (int q = 0; q < RD - 3; ++q) issue(ringmem + q * SLOT);       // steps T-1 .. T-(RD-3)

    // ---- fused zero fill (ZF): the dense gradient plane needs no memset ----------------------------
    // This CTA's rows of step t are one contiguous run (8 B aligned: ncol is even).  Zeroing it
    // with STG costs the issuing warp ~20 cycles per store instruction (7 per step: +60 us on
    // BASELINE config 2, measured) — so lane 0 writes it with ONE TMA bulk store of a zeroed
    // shared-memory buffer (cp.async.bulk.global.shared::cta, SASS UBLKCP.G.S; an 8-byte head /
    // tail where the run is not 16 B aligned), ZD iterations before the sweep stores gradients
    // into that row; cp.async.bulk.wait_group orders the two.  The `rows_before` rows in front of
    // gdyn (the no-grad warm-up rows of the caller's plane) are zeroed the same way, paced over
    // the sweep.
    constexpr int ZD = 16;
    const int znb = ZF ? min(PBPB, d.B - b0w) * d.dyn_ncol : 0;           // floats in this CTA's run
    float* const zbuf = ringmem + RD * SLOT;                                 // znb floats (+ pad), zeroed
    const unsigned zbuf_s = (unsigned)__cvta_generic_to_shared(zbuf);
    int wz_done = 0;                                                         // warm-up rows zeroed so far
    auto zero_run = [&](int64_t row) {      // lane 0 only; row relative to gdyn (negative: rows before)
        char* a = reinterpret_cast<char*>(io.gdyn + (row * d.B + b0w) * d.dyn_ncol);
        int n = znb * 4;
        if (reinterpret_cast<uintptr_t>(a) & 8) { *reinterpret_cast<float2*>(a) = make_float2(0.f, 0.f); a += 8; n -= 8; }
        const int body = n & ~15;
        if (body > 0)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a), "r"(zbuf_s), "r"(body) : "memory");
        if (n & 8) *reinterpret_cast<float2*>(a + body) = make_float2(0.f, 0.f);
    };
    if constexpr (ZF) {
        for (int e = tid; e < ((znb + 3) >> 2) + 1; e += PBPB * PNM) reinterpret_cast<float4*>(zbuf)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (tid == 0) {
            // rows T-2 .. T-ZD: their gradients are stored within the first ZD iterations
            // (row T-1 is never zeroed: it receives the static-parameter and routing gradients)
            // (always ZD - 1 groups, empty where the row does not exist: the wait below counts groups)
            for (int z = T - 2; z >= T - ZD; --z) {
                if (z >= 0) zero_run(z);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }

    // ---- values that cross an iteration boundary ------------------------------------------------
    float rn1 = 0.f, ts1 = 0.f;                      // snow_fwd(t) -> soil_fwd(t): RAIN, tosoil
    float rech1 = 0.f, exc1 = 0.f, slza1 = 0.f;      // soil_fwd(t) -> resp_fwd(t)
    float pet1 = 0.f;                                // PET of the step whose adjoint runs next
    float h1[ND], h2[ND], hd1[ND], hd2[ND];          // dynamic parameters (value, d/d raw) descaled 1 / 2
#pragma unroll                                       // iterations ago (soil: 1, snow: 2 are read)
    for (int s = 0; s < ND; ++s) { h1[s] = h2[s] = hd1[s] = hd2[s] = 0.f; }
    Tape tA, tB, tC;

    // one iteration.  t0: tape of step i (resp fields written here, then consumed), t1: tape of
    // step i-1 (soil fields written here), t2: tape of step i-2 (snow fields written here).
    auto body = [&](auto Uc, Tape& t0, Tape& t1, Tape& t2, const int i, const bool a_sn, const bool a_so,
                    const bool a_rs) {
        constexpr int U = decltype(Uc)::value;
        cp_async_wait<RD - 4>();
        __syncwarp();                // every lane's copies of step i-2 have landed, and every lane is
        issue(ringmem + rb_prev + U * SLOT);      // done with the slot of step i+1, refilled now
        const float* r_sn = ringmem + rb + U * SLOT;
        const float* r_so = ringmem + (U >= 1 ? rb + (U - 1) * SLOT : rb_prev + 2 * SLOT);
        const float* r_rs = ringmem + (U == 2 ? rb : rb_prev + (U + 1) * SLOT);
        if constexpr (ZF) {
            if (a_rs) {
                if (tid == 0) {
                    const int z = i - ZD;                       // the row whose gradients come ZD iterations from now
                    if (z >= 0 && z < T - 1 && z < T - ZD) zero_run(z);
                    const int n_it = T - 1 - i;                 // iterations done so far
                    while ((int64_t)wz_done * T < (int64_t)(n_it + 1) * io.rows_before) {
                        zero_run(-(int64_t)io.rows_before + wz_done);
                        ++wz_done;
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    // the group that zeroed row i was committed ZD groups ago: wait for its writes
                    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(ZD) : "memory");
                }
                __syncwarp();
            }
        }
        float dc[ND], ddc[ND];       // this iteration's descaled dynamic parameters, by stage
        float pf[NPAR];              // parameters as the three forward stages of this iteration see them
#pragma unroll
        for (int k = 0; k < NPAR; ++k) {
            pf[k] = p[k];
            if (DS::is_dyn(k, 0)) {
                const float* rs_ = par_stage(k) == 0 ? r_sn : (par_stage(k) == 1 ? r_so : r_rs);
                pipe_descale_both<SIG>(k, rs_[PARB + DS::slot(k) * 32 + tid], dspan[DS::slot(k)], dlo[DS::slot(k)],
                                       dc[DS::slot(k)], ddc[DS::slot(k)]);
                pf[k] = dc[DS::slot(k)];
            }
        }
        // ---- snow_fwd(i - 2)
        float rn0 = 0.f, ts0 = 0.f;
        if (a_sn) {
            const float4 f = *reinterpret_cast<const float4*>(r_sn + 4 * bl);
            float P = f.x;
            if constexpr (TR::HOURLY) P = P * d.inv_dt;
            float SP = r_sn[STB + 0 * 32 + tid], MW = r_sn[STB + 1 * 32 + tid];
            snow_fwd<VAR, true>(SP, MW, pf, P, f.y, lc, rn0, ts0, t2);
        }
        // ---- soil_fwd(i - 1)
        SoilOut so;
        so.recharge = 0.f; so.excess = 0.f; so.ET = 0.f; so.ef = 0.f; so.capillary = 0.f; so.IE = 0.f;
        float slza0 = 0.f, pet0 = 0.f;
        if (a_so) {
            pet0 = r_so[4 * bl + 2];
            if constexpr (TR::HOURLY) pet0 = pet0 * d.inv_dt;
            float SM = r_so[STB + 2 * 32 + tid];
            slza0 = r_so[STB + 4 * 32 + tid];
            soil_fwd<VAR, BETAET, true>(SM, slza0, pf, rn1, ts1, pet0, lc, so, t1);
        }
        // ---- step i: resp_fwd, then the whole adjoint of the step
        if (a_rs) {
            float pa[NPAR];          // every parameter at step i
#pragma unroll
            for (int k = 0; k < NPAR; ++k) {
                pa[k] = p[k];
                if (DS::is_dyn(k, 0))
                    pa[k] = par_stage(k) == 2 ? dc[DS::slot(k)] : (par_stage(k) == 1 ? h1[DS::slot(k)] : h2[DS::slot(k)]);
            }
            float SUZ = r_rs[STB + 3 * 32 + tid];
            float SLZ = slza1;
            RespOut ro;
            resp_fwd<VAR, true>(SUZ, SLZ, pa, rech1, exc1, lc, ro, t0);
            float gF[HBV_MAX_FLUX];
#pragma unroll
            for (int f = 0; f < HBV_MAX_FLUX; ++f) gF[f] = 0.f;
            gF[HBV_F_QSIM] = r_rs[4 * bl + 3] * inv_nmul;
            float gp[NPAR];      // (time-invariant parameters: terms go straight into the running sum)
#pragma unroll
            for (int k = 0; k < NPAR; ++k) gp[k] = DS::is_dyn(k, 0) ? -0.f : gacc[k];    // (-0 + x folds to x, +0 + x does not)
            float gRE, gW, gPET, gP, gT;
            resp_bwd<VAR, true>(gS[3], gS[4], gF, pa, lc, t0, gp, gRE);       // (cotangent on the streamflow series only)
            soil_bwd<VAR, BETAET, true>(gS[2], gS[4], gRE, gF, pa, pet1, lc, t0, gp, gW, gPET);
            snow_bwd<VAR, true>(gS[0], gS[1], gW, gF, pa, lc, t0, gp, gP, gT);
#pragma unroll
            for (int k = 0; k < NPAR; ++k) {
                if (DS::is_dyn(k, 0)) {
                    const float dd = par_stage(k) == 2 ? ddc[DS::slot(k)] : (par_stage(k) == 1 ? hd1[DS::slot(k)] : hd2[DS::slot(k)]);
                    if (valid) pg[pipe_col<NPAR, DM, LAYOUT>(k)] = gp[k] * dd;
                } else {
                    gacc[k] = gp[k];
                }
            }
            pg -= sd;
        }
        // ---- hand over to the next iteration
        rn1 = rn0; ts1 = ts0;
        rech1 = so.recharge; exc1 = so.excess; slza1 = slza0; pet1 = pet0;
#pragma unroll
        for (int s = 0; s < ND; ++s) { h2[s] = h1[s]; h1[s] = dc[s]; hd2[s] = hd1[s]; hd1[s] = ddc[s]; }
    };

    // iterations i = T+1 .. 0 in trips of three (tape buffers rotate A -> B -> C); padding
    // iterations below 0 have every stage off
    int i = T + 1;
    auto next_trip = [&]() {
        rb_prev = rb;
        rb = (rb == 9 * SLOT) ? 0 : rb + 3 * SLOT;
    };
    auto slow_iter = [&](auto Uc, Tape& t0, Tape& t1, Tape& t2) {
        body(Uc, t0, t1, t2, i, i - 2 >= 0 && i - 2 <= T - 1, i - 1 >= 0 && i - 1 <= T - 1, i >= 0 && i <= T - 1);
        --i;
    };
    auto slow_trip = [&]() {
        slow_iter(IC<0>{}, tA, tB, tC); slow_iter(IC<1>{}, tB, tC, tA); slow_iter(IC<2>{}, tC, tA, tB);
        next_trip();
    };
    slow_trip();                                  // i = T+1, T, T-1
#pragma unroll 1
    while (i - 4 >= 0) {                          // every stage active in all three iterations
        body(IC<0>{}, tA, tB, tC, i, true, true, true); --i;
        body(IC<1>{}, tB, tC, tA, i, true, true, true); --i;
        body(IC<2>{}, tC, tA, tB, i, true, true, true); --i;
        next_trip();
    }
#pragma unroll 1
    while (i >= 0) slow_trip();
    cp_async_wait<0>();
    if constexpr (ZF) {
        if (tid == 0) {
            while (wz_done < io.rows_before) { zero_run(-(int64_t)io.rows_before + wz_done); ++wz_done; }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    }

    // static parameters: d(par)/d(raw) recomputed here, written once (as in hbv_bwd.cu)
    float dps[NPAR];
    uint32_t lastmask = 0;
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, dps, &lastmask);
    if (valid) {
        float* glast = io.gdyn + row_last * d.dyn_ncol + j;
#pragma unroll
        for (int k = 0; k < NPAR; ++k) {
            if (k < d.n_par && !DS::is_dyn(k, 0)) {
                if (lastmask & (1u << k)) glast[d.col[k]] = gacc[k] * dps[k];
                else if (io.gsta != nullptr) io.gsta[(int64_t)b * d.sta_ncol + d.col[k] + j] = gacc[k] * dps[k];
            }
        }
        if (io.gstate_in != nullptr) {
#pragma unroll
            for (int s = 0; s < 5; ++s) io.gstate_in[s * nlane + lane] = gS[s];
        }
    }
}

// ================================================================================================
// host: eligibility + launch
// ================================================================================================
template <int NPAR, int DM, int LAYOUT>
static bool pipe_layout_matches(const KDesc& d) {
    for (int i = 0; i < NPAR; ++i)
        if ((DM >> i) & 1)
            if (d.col[i] != pipe_col<NPAR, DM, LAYOUT>(i)) return false;
    return true;
}

static bool pipe_common_ok(const KDesc& d) {
    if (opt(OPT_PIPE) == 0 || opt(OPT_LEAN) == 0) return false;
    if (d.nmul != PNM || d.nvar != 3 || d.i_prcp != 0 || d.i_tmean != 1 || d.i_pet != 2) return false;
    if (d.T < 2 || d.ck_layout != 0) return false;      // (stored states: plane layout only)
    // Used while the grid leaves schedulers idle (at most one warp per scheduler): there a step's
    // latency IS the kernel time and the skew pays (C2, 531 basins: warm-up 57 -> 44 us, K1 155 ->
    // 137, K2 213 -> 171).  At 2.1 warps per scheduler (BASELINE config 4's per-GPU share, 2,500
    // units) K2s is already issue-bound (80 % of issue slots busy, profiles/r02_ncu_c4.md) and the
    // pipelined forms — more instructions and registers — are slower (fwd 17.0 vs 13.0 ms).
    const long long max_lanes = opt(OPT_PIPE_MAX) >= 0 ? opt(OPT_PIPE_MAX) : 148LL * 4 * 32;
    return (long long)d.B * PNM <= max_lanes;
}

// One-warp CTAs that each run the whole time axis: a grid that does not fit the SMs in ONE wave
// would run its tail after the first wave — twice the time.  Resident CTAs per SM depend on the
// kernel's registers / shared memory, so ask the runtime (once per kernel and device).
template <auto KERN>
static bool pipe_fits_one_wave(size_t smem, int grid) {
    static std::atomic<int> cached[HBV_MAX_DEVICES];      // (one instance per kernel)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= HBV_MAX_DEVICES) return false;
    int per_dev = cached[dev].load(std::memory_order_acquire);
    if (per_dev == 0) {
        int occ = 0, sms = 0;
        // many small CTAs per SM: ask for the largest shared-memory carve-out, or the default
        // split caps the resident CTAs well below what the registers allow
        cudaFuncSetAttribute(KERN, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERN, PBPB * PNM, smem) != cudaSuccess) return false;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
        per_dev = occ * sms;
        if (per_dev <= 0) return false;
        cached[dev].store(per_dev, std::memory_order_release);
    }
    return grid <= per_dev;
}

template <auto KERN, typename IO>
static int pipe_launch(const KDesc& d, const IO& io, size_t smem, cudaStream_t st) {
    const int grid = (d.B + PBPB - 1) / PBPB;
    if (!pipe_fits_one_wave<KERN>(smem, grid)) return HBV_NOT_ELIGIBLE;
    KERN<<<grid, PBPB * PNM, smem, st>>>(d, io);
    count_launch();
    count_lean_launch();
    count_pipe_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

constexpr int PRD_F = 12;    // forward ring depth (steps)
constexpr int PRD_B = 12;    // adjoint ring depth

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG>
static int launch_fwd_pipe(KDesc d, const FwdPtrs& io, cudaStream_t st) {
    constexpr int ND = DynSet<Traits<VAR>::NPAR, DM>::NDYN;
    d.BPB = PBPB;
    const size_t smem = ((size_t)2 * PTC * PBPB * (PNM * NFP + 12) + (size_t)PRD_F * (8 + 32 * ND)) * sizeof(float);
    if (io.ckpt != nullptr)
        return pipe_launch<hbv_fwd_pipe_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, true, PRD_F>>(d, io, smem, st);
    return pipe_launch<hbv_fwd_pipe_kernel<VAR, BETAET, DM, LAYOUT, SIG, false, true, PRD_F>>(d, io, smem, st);
}

template <int VAR, bool BETAET>
int try_fwd_pipe_warm(const KDesc& d0, const FwdPtrs& io, cudaStream_t st) {
    if (!pipe_common_ok(d0)) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.state_series != nullptr || io.ckpt != nullptr)
        return HBV_NOT_ELIGIBLE;
    KDesc d = d0;
    d.BPB = PBPB;
    const size_t smem = (size_t)PRD_F * 8 * sizeof(float);
    if (d.apply_sigmoid)
        return pipe_launch<hbv_fwd_pipe_kernel<VAR, BETAET, 0, 0, true, false, false, PRD_F>>(d, io, smem, st);
    return pipe_launch<hbv_fwd_pipe_kernel<VAR, BETAET, 0, 0, false, false, false, PRD_F>>(d, io, smem, st);
}

template <int VAR, bool BETAET, int DM>
int try_fwd_pipe(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if (!write_flux || !pipe_common_ok(d)) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.state_series != nullptr) return HBV_NOT_ELIGIBLE;
    if (io.ckpt != nullptr && d.K != 1) return HBV_NOT_ELIGIBLE;
    if (d.dyn_ncol % 2 != 0 || reinterpret_cast<uintptr_t>(io.dyn) % 8 != 0) return HBV_NOT_ELIGIBLE;   // 8 B ring copies
    for (int f = 0; f < Traits<VAR>::NFLUX; ++f)
        if (io.flux[f] == nullptr) return HBV_NOT_ELIGIBLE;
    const bool sig = d.apply_sigmoid != 0;
    if (pipe_layout_matches<NPAR, DM, 0>(d))
        return sig ? launch_fwd_pipe<VAR, BETAET, DM, 0, true>(d, io, st) : launch_fwd_pipe<VAR, BETAET, DM, 0, false>(d, io, st);
    if (pipe_layout_matches<NPAR, DM, 1>(d))
        return sig ? launch_fwd_pipe<VAR, BETAET, DM, 1, true>(d, io, st) : launch_fwd_pipe<VAR, BETAET, DM, 1, false>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG>
static int launch_bwd_pipe(KDesc d, const BwdPtrs& io, cudaStream_t st) {
    constexpr int ND = DynSet<Traits<VAR>::NPAR, DM>::NDYN;
    d.BPB = PBPB;
    const size_t smem = (size_t)PRD_B * (8 + 32 * (ND + 5)) * sizeof(float);
    if (io.zero_fill && (popc_c((unsigned)DM) * PNM != d.dyn_ncol || io.rows_before > 0)) {
        const size_t zbytes = ((size_t)PBPB * d.dyn_ncol * sizeof(float) + 31) / 16 * 16;    // the zeroed source buffer
        return pipe_launch<hbv_bwd_pipe_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, PRD_B>>(d, io, smem + zbytes, st);
    }
    if (io.rows_before > 0) return HBV_NOT_ELIGIBLE;
    return pipe_launch<hbv_bwd_pipe_kernel<VAR, BETAET, DM, LAYOUT, SIG, false, PRD_B>>(d, io, smem, st);
}

template <int VAR, bool BETAET, int DM>
int try_bwd_pipe(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if (d.K != 1 || !pipe_common_ok(d) || d.T < 3) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.gmuwts != nullptr || io.gforcing != nullptr ||
        io.gstate_series != nullptr || io.gdyn == nullptr) return HBV_NOT_ELIGIBLE;
    if (d.dyn_ncol % 2 != 0 || reinterpret_cast<uintptr_t>(io.dyn) % 8 != 0 ||
        reinterpret_cast<uintptr_t>(io.ckpt) % 16 != 0 || reinterpret_cast<uintptr_t>(io.gdyn) % 8 != 0)
        return HBV_NOT_ELIGIBLE;
    if (io.gflux[HBV_F_QSIM] == nullptr) return HBV_NOT_ELIGIBLE;
    for (int f = 1; f < HBV_MAX_FLUX; ++f)
        if (io.gflux[f] != nullptr) return HBV_NOT_ELIGIBLE;
    const bool sig = d.apply_sigmoid != 0;
    if (pipe_layout_matches<NPAR, DM, 0>(d))
        return sig ? launch_bwd_pipe<VAR, BETAET, DM, 0, true>(d, io, st) : launch_bwd_pipe<VAR, BETAET, DM, 0, false>(d, io, st);
    if (pipe_layout_matches<NPAR, DM, 1>(d))
        return sig ? launch_bwd_pipe<VAR, BETAET, DM, 1, true>(d, io, st) : launch_bwd_pipe<VAR, BETAET, DM, 1, false>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

// the compiled (variant, dynamic set) pairs: those of hbv_lean.cu
#if HBV_IN_PART(1)
template int try_fwd_pipe<HBV_VARIANT_HBV, true, DM_D2>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_pipe<HBV_VARIANT_HBV11P, true, DM_D2>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_pipe<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_pipe<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_pipe_warm<HBV_VARIANT_HBV, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_pipe_warm<HBV_VARIANT_HBV, false>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_pipe_warm<HBV_VARIANT_HBV11P, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_pipe_warm<HBV_VARIANT_HBV2, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_pipe_warm<HBV_VARIANT_HOURLY, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
#endif
#if HBV_IN_PART(2)
template int try_bwd_pipe<HBV_VARIANT_HBV, true, DM_D2>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_pipe<HBV_VARIANT_HBV11P, true, DM_D2>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_pipe<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_pipe<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);
#endif

}  // namespace hbv
