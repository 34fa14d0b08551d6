// hbv_lean.cu — K1s / K2s: K1 (hbv_fwd.cu) and the every-state sweep of K2 (hbv_bwd.cu) compiled
// for the layout the reference actually produces, in the throughput regime (large grids).
//
// ncu on the north-star shard (22,500 basins, `hbv` with the shipped dynamic set
// [parBETA, parBETAET]; profiles/r01_ncu_shard_d2.md) showed both kernels bound by instruction
// issue, and more than half of the issued instructions were not HBV arithmetic: 64-bit address
// arithmetic re-derived every step from runtime column numbers and strides (5 instructions per
// 4-byte load, 10 per gradient store, 23 for the five stored-state loads), the generic sweep
// schedule, per-step tests of options nobody had switched on (muwts, state series, forcing
// gradients, per-series upstream gradients, zero fill), descale constants re-read from the
// constant bank.  The kernels here take all of that out of the loop for the common case:
//   * forcing columns (prcp, tmean, pet) = (0, 1, 2) of a 3-wide x_phy, nmul = 16, parameter i at
//     column 16*i (packed form, hbv.py:201-208) or dynamic parameter s at 16*s (split form,
//     hbv_2.py:211-230): every load / store is `[running pointer + immediate]`;
//   * five running pointers advanced by a constant stride per step (forcing row, parameter row,
//     gradient row, upstream-gradient row, stored states), the stored states walked as ONE
//     pointer because consecutive (t, state) planes are equidistant;
//   * a compile-time dynamic set, compile-time sigmoid, descale constants in registers;
//   * adjoint: upstream gradient on the streamflow series only (the training loss), so the other
//     eleven series' adjoint terms fold away at compile time.
// Anything else (dropout masks, muwts, state series, other cotangents, other layouts, K > 1)
// takes K1 / K2.  Arithmetic is the same hbv_step.cuh code: results agree with K1 / K2 to fp32
// contraction noise (tests/test_lean_gpu.py).
//
// Reference spans replaced: models/hbv/hbv.py:423-511, hbv_1_1p.py:422-524, hbv_2.py:464-585,
// hbv_2_hourly.py:527-683 and PyTorch autograd over them.
#include <atomic>
#include <cstdlib>
#include "hbv_common.cuh"

namespace hbv {

constexpr int LNM = 16;      // components per basin
// basins per CTA (template LBPB): 8 (128 threads) on large grids; 2 (one warp, so the chunk barrier
// costs nothing and the CTAs spread over all SMs) on small, latency-bound ones
constexpr int LTC = 4;       // forward: time steps per output chunk
constexpr int LNCH = 4;      // forward, chunk-ring form (RD = -1): chunks in the ring

template <int NPAR, int DM, int LAYOUT>
__host__ __device__ constexpr int lean_col(int i) {
    return (LAYOUT == 0 ? i : DynSet<NPAR, DM>::slot(i)) * LNM;
}

template <bool SIG>
__device__ __forceinline__ float lean_descale(int i, float raw, float span, float lo) {
    if constexpr (SIG) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        return fmaf(s, span, lo);
    } else {
        return fmaf(raw, span, lo);
    }
}

template <bool SIG>
__device__ __forceinline__ void lean_descale_both(int i, float raw, float span, float lo, float& v, float& dv) {
    if constexpr (SIG) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        v = fmaf(s, span, lo);
        dv = span * s * (1.0f - s);
    } else {
        v = fmaf(raw, span, lo);
        dv = span;
    }
}

// ================================================================================================
// K1s: forward.  CK: store the state before every step (K = 1) for the adjoint.
// ================================================================================================
// RD: 0 = inputs prefetched into registers (large grids: many resident warps hide HBM latency);
// > 0 = one-warp CTA with a shared-memory ring of RD steps filled with cp.async (small grids: one
// warp owns a scheduler, only prefetch DISTANCE hides latency, and register loads cannot provide
// it — loads in flight share the warp's six scoreboard slots, so waiting for the oldest waits for
// younger ones too; ncu showed 55-64 % of all stall samples on the first use of a prefetched
// register)
// WF: write the flux planes (false: the warm-up / `initialize` run — states only, hbv.py:327-346)
// KS: the state before a step is stored every KS-th step (KS = 1, 2 or 4; LTC % KS == 0, so
// which steps of an output chunk store is known at compile time)
// RD = -1 (128-thread CTAs): the same cp.async staging at the granularity of an output chunk — a
// ring of LNCH chunks of LTC steps, one commit group per chunk, filled two to three chunks (8-12
// steps) ahead; the wait for a chunk rides on the barrier that already closes the previous chunk's
// nmul reduction.  Why not registers: ptxas puts every LDG of the loop on ONE scoreboard
// (cuobjdump control codes: write barrier SB5 on all 24 loads of the four-buffer register form), so
// the first use of the oldest buffer also waits for the loads issued a moment ago — ncu attributed
// 28 % of the kernel's stall samples to that single wait on BASELINE config 4's per-GPU grid
// (2.1 warps per scheduler).  cp.async groups complete in FIFO order and are waited for by count.
// CKL: layout of the state store (hbv_common.cuh) as a compile-time fact — warp-major puts the five
// states of a stored step at immediate offsets from one pointer (10 instructions per step less than
// five planes a runtime stride apart); compiled for the chunk-ring form with KS 1 / 4 only
template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, bool CK, int LBPB, int RD, bool WF = true, int KS = 1,
          int CKL = 0>
__global__ void __launch_bounds__(LBPB * LNM, LBPB == 8 ? (RD < 0 ? 4 : 6) : 1)
hbv_fwd_lean_kernel(const KDesc d, const FwdPtrs io) {
    static_assert(LTC % KS == 0, "checkpoint interval must divide the output chunk");
    static_assert(RD >= 0 || WF, "the chunk ring rides on the barriers of the flux reduction");
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    constexpr int ND = DS::NDYN > 0 ? DS::NDYN : 1;     // (array extent; no slot is used when DM = 0)
    extern __shared__ __align__(16) float tile[];

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b_raw = blockIdx.x * LBPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * LNM + j;
    const int64_t nlane = (int64_t)d.B * LNM;

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

    // running pointers (advanced by a constant stride per step)
    const float* pf = io.forcing + (int64_t)b * 3;
    const float* pd = io.dyn + (int64_t)b * d.dyn_ncol + j;
    const int64_t sf = (int64_t)d.B * 3, sd = (int64_t)d.B * d.dyn_ncol;
    // consecutive (segment, state) planes are ck_plane apart in either layout of the store
    // (hbv_common.cuh: planes over all lanes, or warp-major — 640 contiguous bytes per warp and
    // stored step; with the stores kept in L2 altogether this kernel takes 7.27 instead of 8.02 ms
    // on BASELINE config 4's per-GPU grid, with the warp-major layout 7.82)
    float* pk = CK ? io.ckpt + (CKL ? (lane >> 5) * (ck_nseg(d) * 160) + (lane & 31) : lane) : nullptr;

    struct In { float P, T, E; float raw[ND]; };
    int t_issue = 0;
    auto load = [&](In& in) {
        in.P = __ldg(pf); in.T = __ldg(pf + 1); in.E = __ldg(pf + 2);
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0)) in.raw[DS::slot(i)] = __ldg(pd + lean_col<NPAR, DM, LAYOUT>(i));
        if (++t_issue < d.T) { pf += sf; pd += sd; }   // the last row is simply loaded again
    };

    // output staging tile + this thread's reduce item (hbv_fwd.cu)
    constexpr int bstride = LNM * NFP + 12;
    constexpr int tstride_s = LBPB * bstride;
    float* const my_slot = tile + bl * bstride + j * NFP;
    constexpr float inv_nmul = 1.0f / (float)LNM;
    constexpr int items = LTC * LBPB * 3;            // 96 <= 128 threads (24 <= 32)
    const int r_q = tid % 3;
    const int r_r = tid / 3;
    const int r_bl = r_r % LBPB;
    const int r_tc = r_r / LBPB;
    const int r_bb = blockIdx.x * LBPB + r_bl;
    const bool r_ok = (tid < items) && (r_bb < d.B);
    const float* const r_src = tile + r_tc * tstride_s + r_bl * bstride + r_q * 4;
    int64_t r_o = (int64_t)r_tc * d.B + (r_ok ? r_bb : 0);     // this item's element of its four planes
    const int64_t r_adv = (int64_t)LTC * d.B;

    Tape tp;
    auto do_step = [&](const In& in, int tc) {
        if constexpr (CK) {
            if (tc % KS == 0) {        // (tc is a literal at every call site)
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    if constexpr (CKL) { if (valid) pk[s * 32] = S[s]; }
                    else { if (valid) *pk = S[s]; pk += nlane; }
                }
                if constexpr (CKL) pk += 160;
            }
        }
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0)) p[i] = lean_descale<SIG>(i, in.raw[DS::slot(i)], dspan[DS::slot(i)], dlo[DS::slot(i)]);
        float P = in.P, PET = in.E;
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float F[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, false>(S, p, P, in.T, PET, lc, F, tp);
        if constexpr (WF) {
            float4* o4 = reinterpret_cast<float4*>(my_slot + tc * tstride_s);
            o4[0] = make_float4(F[0], F[1], F[2], F[3]);
            o4[1] = make_float4(F[4], F[5], F[6], F[7]);
            o4[2] = make_float4(F[8], F[9], F[10], TR::NFLUX > 11 ? F[11] : 0.f);
        }
    };

    // ring (RD != 0) after the output tile: the CTA's inputs of a step are staged by cp.async
    // (see the adjoint below for the why):
    //   A  4 B x 3 LBPB lanes   P, T, PET of the CTA's basins            -> slot[4 bl + k]
    //   B  8 B x 8 LBPB ND      the 64 B run of every dynamic parameter   -> slot[PARB + NT k + tid]
    // read back with one LDS.128 + ND LDS; a lane reads what other lanes copied, so the wait for a
    // group is followed by __syncwarp() (one-warp CTA, RD > 0: one group per step) or rides on the
    // chunk barrier (RD < 0: one group per output chunk).
    constexpr int NT = LBPB * LNM;
    constexpr int NDR = DS::NDYN;
    constexpr int PARB = 4 * LBPB, SLOT = PARB + NT * NDR;
    constexpr int NB = (8 * LBPB * NDR + NT - 1) / NT;
    constexpr int RDS = RD > 0 ? RD : (RD < 0 ? LNCH * LTC : 2);
    float* const ring0 = tile + (WF ? LTC * tstride_s : 0);
    float* const ring_end = ring0 + RDS * SLOT;
    float* wp = ring0;
    const float* rp = ring0;
    const int b0w = blockIdx.x * LBPB;
    const int kA = tid & 15, bbA = tid >> 4;
    const bool actA = kA < 3;
    // sources as (row-0 pointer, byte stride per step): the address of step t is one IMAD.WIDE.U32
    // (base + t * stride) instead of a 64-bit running pointer per source plus the moves into the
    // even register pair LDGSTS wants.  Forward only: measured on B200 this takes K1p at C2 from
    // 134 to 117 us, while the same change in the adjoints (five sources) removed 11-15 instructions
    // per step and made them no faster (K2s at config 4: 11.65 -> 11.70 ms, K2p at C2: 161 -> 170 us;
    // IMAD.WIDE is not a full-rate instruction).
    const char* const baseA = reinterpret_cast<const char*>(io.forcing + (int64_t)min(b0w + bbA, d.B - 1) * 3 + kA);
    const uint32_t strA = (uint32_t)d.B * 12u, strB = (uint32_t)d.B * (uint32_t)d.dyn_ncol * 4u;
    const int dstA = 4 * bbA + kA;
    const char* baseB[NB > 0 ? NB : 1];
    int dstB[NB > 0 ? NB : 1];
    bool actB[NB > 0 ? NB : 1];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        const int gq = o * NT + tid;
        actB[o] = gq < 8 * LBPB * NDR;
        const int r = actB[o] ? (gq >> 3) : 0, q = gq & 7;
        const int bb = r / (NDR > 0 ? NDR : 1), k = r - bb * NDR;
        int col = 0;
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0) && DS::slot(i) == k) col = lean_col<NPAR, DM, LAYOUT>(i);
        baseB[o] = reinterpret_cast<const char*>(io.dyn + (int64_t)min(b0w + bb, d.B - 1) * d.dyn_ncol + col + 2 * q);
        dstB[o] = PARB + k * NT + bb * 16 + 2 * q;
    }
    auto stage = [&](float* w) {     // copies of the next time step (the last row again past the end)
        const uint64_t ts = (uint32_t)t_issue;
        if (actA) cp_async4(w + dstA, reinterpret_cast<const float*>(baseA + ts * strA));
#pragma unroll
        for (int o = 0; o < NB; ++o)
            if (actB[o]) cp_async8(w + dstB[o], reinterpret_cast<const float*>(baseB[o] + ts * strB));
        if (t_issue + 1 < d.T) ++t_issue;
    };
    auto issue = [&]() {             // one-warp form: one step, one group
        stage(wp);
        cp_async_commit();
        wp += SLOT;
        if (wp == ring_end) wp = ring0;
    };
    auto issue_chunk = [&]() {       // chunk form: LTC steps, one group
#pragma unroll
        for (int u = 0; u < LTC; ++u) stage(wp + u * SLOT);
        cp_async_commit();
        wp += LTC * SLOT;
        if (wp == ring_end) wp = ring0;
    };
    auto fetch = [&](In& in, const float* r) {
        const float4 f = *reinterpret_cast<const float4*>(r + 4 * bl);
        in.P = f.x; in.T = f.y; in.E = f.z;
#pragma unroll
        for (int k = 0; k < NDR; ++k) in.raw[k] = r[PARB + k * NT + tid];
    };
    auto pop = [&](In& in) {         // oldest staged step
        cp_async_wait<(RD > 0 ? RD : 2) - 2>();
        __syncwarp();
        issue();
        fetch(in, rp);
        rp += SLOT;
        if (rp == ring_end) rp = ring0;
    };

    // registers (RD == 0): two named prefetch buffers of two steps each (see hbv_fwd.cu: loads in
    // flight must not share a scoreboard slot with the values being consumed)
    In A0, A1, B0, B1;
    if constexpr (RD > 0) {
#pragma unroll 1
        for (int q = 0; q < RDS - 1; ++q) issue();
    } else if constexpr (RD == 0) {
        load(A0); load(A1);
    } else {
#pragma unroll 1
        for (int q = 0; q < LNCH - 1; ++q) issue_chunk();
        cp_async_wait<LNCH - 2>();
        __syncthreads();             // chunk 0 has landed for every thread
    }
    auto reduce_out = [&](int tcn) {          // nmul reduction of the chunk's tile -> [T, B] planes
        if constexpr (WF) {
            __syncthreads();
            if (r_ok && r_tc < tcn) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int jj = 0; jj < LNM; ++jj) {
                    const float4 v = *reinterpret_cast<const float4*>(r_src + jj * NFP);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                float* const* pl = io.flux + r_q * 4;
                pl[0][r_o] = acc.x * inv_nmul;
                pl[1][r_o] = acc.y * inv_nmul;
                pl[2][r_o] = acc.z * inv_nmul;
                if (r_q * 4 + 3 < TR::NFLUX) pl[3][r_o] = acc.w * inv_nmul;
            }
            r_o += r_adv;
            // chunk form: this thread's copies of the NEXT chunk have landed; the barrier makes
            // that true for every thread's (groups in flight after it: LNCH - 2)
            if constexpr (RD < 0) cp_async_wait<LNCH - 2>();
            __syncthreads();
        }
    };
    for (int t0 = 0; t0 < d.T; t0 += LTC) {
        const int tcn = min(LTC, d.T - t0);
        if constexpr (RD > 0) {
#pragma unroll
            for (int u = 0; u < LTC; ++u)
                if (u < tcn) { pop(A0); do_step(A0, u); }
        } else if constexpr (RD < 0) {
            issue_chunk();           // into the slot of the chunk every thread finished before the last barrier
#pragma unroll
            for (int u = 0; u < LTC; ++u)
                if (u < tcn) { fetch(A0, rp + u * SLOT); do_step(A0, u); }
            rp += LTC * SLOT;
            if (rp == ring_end) rp = ring0;
        } else {
            load(B0); load(B1);
            do_step(A0, 0);
            if (1 < tcn) do_step(A1, 1);
            load(A0); load(A1);
            if (2 < tcn) do_step(B0, 2);
            if (3 < tcn) do_step(B1, 3);
        }
        reduce_out(tcn);
    }
    if constexpr (RD != 0) cp_async_wait<0>();
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

// ================================================================================================
// K2s: adjoint, every state stored (K = 1), upstream gradient on the streamflow series
// ================================================================================================
// PF: input prefetch distance in time steps (register buffers, the loop is unrolled PF times):
// 1 where many resident warps hide HBM latency, 3 on small grids where one warp owns a scheduler
// RD as in the forward kernel (0: registers, PF steps ahead; > 0: cp.async ring of RD steps)
// ZF (one-warp CTAs only): write every element of the gradient rows 0 .. T-2 — the CTA's rows of a
// step are one contiguous run; the warp zeroes it with 8 B stores, __syncwarp(), then stores the
// gradients on top — so the dense plane needs no memset (hbv_bwd_io_t.gdyn_zero_fill).
// KS (ring form only): the forward stored the state every KS-th step.  The sweep then walks
// segments of KS steps last-to-first: the segment's inputs (KS ring slots) go to registers, the
// KS-1 missing states are recomputed from the stored one and kept in registers, and the KS steps
// are swept in reverse — one extra tape-free forward step per missing state instead of 320 B of
// state traffic per basin-step written by the forward and read back here.
template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, int LBPB, int PF, int RD, bool ZF = false, int KS = 1,
          int CKL = 0>      // CKL: as in the forward kernel (ring form, KS 1 / 4)
// (one-warp form: registers unconstrained — 17 resident warps per SM; forcing 20-28 was measured
// equal or slower on the 22.5k-basin shard, 4.31 / 5.55 / 5.72 ms: the kernel is HBM-bound there)
__global__ void __launch_bounds__(LBPB * LNM, LBPB == 8 ? 4 : 1)
hbv_bwd_lean_kernel(const KDesc d, const BwdPtrs io) {
    static_assert(!ZF || LBPB == 2, "fused zero fill needs a one-warp CTA");
    static_assert(KS == 1 || RD > 0, "the segment sweep is built on the ring form");
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    constexpr int ND = DS::NDYN;

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b_raw = blockIdx.x * LBPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * LNM + j;
    const int64_t nlane = (int64_t)d.B * LNM;

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

    // running pointers, positioned on step T-1 and walked backwards
    const int64_t row_last = (int64_t)(d.T - 1) * d.B + b;
    const int64_t sf = (int64_t)d.B * 3, sd = (int64_t)d.B * d.dyn_ncol;
    const float* pf = io.forcing + row_last * 3;
    const float* pd = io.dyn + row_last * d.dyn_ncol + j;
    const float* pq = io.gflux[HBV_F_QSIM] + row_last;
    static_assert(CKL == 0 || (RD > 0 && LBPB == 2), "warp-major state store: ring form only");
    const float* pc = io.ckpt + (int64_t)d.T * 5 * nlane + lane;     // one plane past (T-1, state 4)
    float* pg = io.gdyn + row_last * d.dyn_ncol + j;
    constexpr float inv_nmul = 1.0f / (float)LNM;

    struct In { float P, T, E, q; float raw[ND]; float S[5]; };
    auto load = [&](In& in) {
        in.P = __ldg(pf); in.T = __ldg(pf + 1); in.E = __ldg(pf + 2);
        in.q = __ldg(pq);
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0)) in.raw[DS::slot(i)] = __ldg(pd + lean_col<NPAR, DM, LAYOUT>(i));
#pragma unroll
        for (int s = 4; s >= 0; --s) { pc -= nlane; in.S[s] = __ldg(pc); }
        pf -= sf; pd -= sd; pq -= d.B;
    };

    // ---- ring form (RD > 0, one-warp CTA): the warp's inputs of a step are staged by FOUR cp.async
    // instructions instead of one per value and lane (LDGSTS costs the LSU ~8 cycles per warp
    // instruction whatever its width; ncu showed the 11-instruction form LSU-limited on large grids):
    //   A  4 B x 8 lanes   P, T, PET and dL/dQ of the two basins         -> slot[0 .. 7]
    //   B  8 B x 16*ND     the 64 B run of every dynamic parameter        -> slot[8 + 32 k + lane]
    //   C 16 B x 32 lanes  stored states 0..3 (128 B per state and warp)  -> slot[ST + 32 s + lane]
    //   D 16 B x 8 lanes   stored state 4
    // and read back by every lane with one LDS.128 (P, T, PET, dL/dQ) + ND + 5 LDS.  A lane now reads
    // what other lanes copied: cp.async.wait_group is followed by __syncwarp().
    extern __shared__ __align__(16) float ringmem[];
    constexpr int NDR = DS::NDYN;
    constexpr int PARB = 8, STB = 8 + 32 * NDR, SLOT = 8 + 32 * (NDR + 5);
    constexpr int NB = (16 * NDR + 31) / 32;
    constexpr int RDS = RD > 0 ? RD : 2;
    float* const ring_end = ringmem + RDS * SLOT;
    float* wp = ringmem;
    const float* rp = ringmem;
    const int b0w = blockIdx.x * LBPB;
    const int nbw = min(LBPB, d.B - b0w);
    const int kA = tid & 15, bbA = tid >> 4;
    const bool actA = kA < 4;
    const int64_t rowA = (int64_t)(d.T - 1) * d.B + min(b0w + bbA, d.B - 1);
    const float* srcA = (kA < 3) ? io.forcing + rowA * 3 + kA : io.gflux[HBV_F_QSIM] + rowA;
    const int64_t strA = (kA < 3) ? sf : (int64_t)d.B;
    const int dstA = 4 * bbA + kA;
    const float* srcB[NB];
    int dstB[NB];
    bool actB[NB];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        const int gq = o * 32 + tid;
        actB[o] = gq < 16 * NDR;
        const int r = actB[o] ? (gq >> 3) : 0, q = gq & 7;
        const int bb = r / (NDR > 0 ? NDR : 1), k = r - bb * NDR;
        int col = 0;
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0) && DS::slot(i) == k) col = lean_col<NPAR, DM, LAYOUT>(i);
        srcB[o] = io.dyn + ((int64_t)(d.T - 1) * d.B + min(b0w + bb, d.B - 1)) * d.dyn_ncol + col + 2 * q;
        dstB[o] = PARB + k * 32 + bb * 16 + 2 * q;
    }
    const int sC = tid >> 3, qC = tid & 7;
    const bool actC = 4 * qC < 16 * nbw;
    // stored states: plane (segment, state) at ckpt + (5 segment + state) nlane, segment = t / KS
    const int64_t seg_last = (d.T - 1) / KS;
    // (warp-major store: a one-warp CTA's 32 lanes are one 32-lane group, b0w * LNM = 32 * blockIdx.x)
    const int64_t ckw = CKL ? (int64_t)blockIdx.x * (ck_nseg(d) * 160) : (int64_t)b0w * LNM;
    const int64_t ckp = CKL ? 32 : nlane;
    const float* srcC = io.ckpt + ckw + (seg_last * 5 + sC) * ckp + 4 * qC;
    const float* srcD = io.ckpt + ckw + (seg_last * 5 + 4) * ckp + 4 * qC;
    const bool actD = actC && tid < 8;
    const int dstC = STB + sC * 32 + 4 * qC, dstD = STB + 4 * 32 + 4 * qC;
    const int64_t strC = CKL ? 160 : 5 * nlane;
    int t_stage = d.T - 1;
    auto issue = [&]() {             // stage the inputs of the next step of the sweep (if any)
        if (t_stage >= 0) {
            if (actA) cp_async4(wp + dstA, srcA);
#pragma unroll
            for (int o = 0; o < NB; ++o)
                if (actB[o]) cp_async8(wp + dstB[o], srcB[o]);
            if (KS == 1 || t_stage % KS == 0) {      // the state before this step was stored
                if (actC) cp_async16(wp + dstC, srcC);
                if (actD) cp_async16(wp + dstD, srcD);
                srcC -= strC; srcD -= strC;
            }
            srcA -= strA;
#pragma unroll
            for (int o = 0; o < NB; ++o) srcB[o] -= sd;
        }
        --t_stage;
        cp_async_commit();
        wp += SLOT;
        if (wp == ring_end) wp = ringmem;
    };
    auto pop = [&](In& in) {
        cp_async_wait<RDS - 2>();
        __syncwarp();                // every lane's copies of the oldest step have landed, and every
        issue();                     // lane is done reading the slot that is refilled now
        const float4 f = *reinterpret_cast<const float4*>(rp + 4 * bl);
        in.P = f.x; in.T = f.y; in.E = f.z; in.q = f.w;
#pragma unroll
        for (int k = 0; k < NDR; ++k) in.raw[k] = rp[PARB + k * 32 + tid];
#pragma unroll
        for (int s2 = 0; s2 < 5; ++s2) in.S[s2] = rp[STB + s2 * 32 + tid];
        rp += SLOT;
        if (rp == ring_end) rp = ringmem;
    };

    Tape tp;
    In buf[PF];
    int t_load = d.T - 1;
    auto load_next = [&](In& in) {
        if (t_load >= 0) load(in);
        --t_load;
    };
    if constexpr (RD > 0) {
#pragma unroll 1
        for (int q = 0; q < RDS - 1; ++q) issue();
    } else {
#pragma unroll
        for (int u = 0; u < PF; ++u) load_next(buf[u]);
    }

    // fused zero fill: this CTA's run of row t as float2 (rows are 8 B aligned: ncol is even)
    const int zb0 = blockIdx.x * LBPB;
    const int nz2 = ZF ? (min(LBPB, d.B - zb0) * d.dyn_ncol) >> 1 : 0;
    float2* pz = ZF ? reinterpret_cast<float2*>(io.gdyn + ((int64_t)(d.T - 1) * d.B + zb0) * d.dyn_ncol) + tid : nullptr;
    const int64_t sd2 = sd >> 1;
    int t_cur = d.T - 1;

    auto process = [&](const In& cur, const float (&S_in)[5]) {
        if constexpr (ZF) {
            if (t_cur < d.T - 1) {      // row T-1 also holds the static-parameter and routing gradients
#pragma unroll 4
                for (int e = tid; e < nz2; e += LBPB * LNM) pz[e - tid] = make_float2(0.f, 0.f);
            }
            __syncwarp();
            pz -= sd2;
            --t_cur;
        }
        float dpd[ND];
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0))
                lean_descale_both<SIG>(i, cur.raw[DS::slot(i)], dspan[DS::slot(i)], dlo[DS::slot(i)], p[i], dpd[DS::slot(i)]);
        float S[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) S[s] = S_in[s];
        float P = cur.P, PET = cur.E;
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float Fl[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, true>(S, p, P, cur.T, PET, lc, Fl, tp);

        float gF[HBV_MAX_FLUX];
#pragma unroll
        for (int f = 0; f < HBV_MAX_FLUX; ++f) gF[f] = 0.f;
        gF[HBV_F_QSIM] = cur.q * inv_nmul;
        // the adjoint stages only ever add to gp[]: a time-invariant parameter's terms go straight
        // into its running sum (one FFMA per term instead of a per-step sum plus an FADD)
        float gp[NPAR];
#pragma unroll
        for (int i = 0; i < NPAR; ++i) gp[i] = DS::is_dyn(i, 0) ? -0.f : gacc[i];     // (-0 + x folds to x, +0 + x does not)
        float gX[3];
        step_bwd<VAR, BETAET, true>(gS, gF, p, PET, lc, tp, gp, gX);      // (cotangent on the streamflow series only)

#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (DS::is_dyn(i, 0)) { if (valid) pg[lean_col<NPAR, DM, LAYOUT>(i)] = gp[i] * dpd[DS::slot(i)]; }
            else gacc[i] = gp[i];
        }
        pg -= sd;
    };
    if constexpr (RD > 0 && KS > 1) {
        // segment sweep: inputs of the segment's steps in registers (seg[k] = step t0 + k), the
        // stored state of step t0 arrives with that step's ring slot
#pragma unroll 1
        for (int t0 = (int)seg_last * KS; t0 >= 0; t0 -= KS) {
            const int len = min(KS, d.T - t0);
            In seg[KS];
#pragma unroll
            for (int k = KS - 1; k >= 0; --k)
                if (k < len) pop(seg[k]);
            float Sg[KS][5];
#pragma unroll
            for (int s = 0; s < 5; ++s) Sg[0][s] = seg[0].S[s];
#pragma unroll
            for (int k = 0; k + 1 < KS; ++k) {
                if (k + 1 < len) {
#pragma unroll
                    for (int i = 0; i < NPAR; ++i)
                        if (DS::is_dyn(i, 0))
                            p[i] = lean_descale<SIG>(i, seg[k].raw[DS::slot(i)], dspan[DS::slot(i)], dlo[DS::slot(i)]);
                    float S[5];
#pragma unroll
                    for (int s = 0; s < 5; ++s) S[s] = Sg[k][s];
                    float P = seg[k].P, PET = seg[k].E;
                    if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
                    float Fl[HBV_MAX_FLUX];
                    step_fwd<VAR, BETAET, false>(S, p, P, seg[k].T, PET, lc, Fl, tp);
#pragma unroll
                    for (int s = 0; s < 5; ++s) Sg[k + 1][s] = S[s];
                }
            }
#pragma unroll
            for (int k = KS - 1; k >= 0; --k)
                if (k < len) process(seg[k], Sg[k]);
        }
        cp_async_wait<0>();
    } else if constexpr (RD > 0) {
#pragma unroll 1
        for (int t = d.T - 1; t >= 0; --t) {
            pop(buf[0]);
            process(buf[0], buf[0].S);
        }
        cp_async_wait<0>();
    } else {
#pragma unroll 1
        for (int t = d.T - 1; t >= 0; t -= PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (t - u >= 0) {
                    const In cur = buf[u];
                    load_next(buf[u]);
                    process(cur, cur.S);
                }
            }
        }
    }

    // static parameters: d(par)/d(raw) recomputed here, written once (as in hbv_bwd.cu)
    float dps[NPAR];
    uint32_t lastmask = 0;
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, dps, &lastmask);
    if (valid) {
        float* glast = io.gdyn + row_last * d.dyn_ncol + j;
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (i < d.n_par && !DS::is_dyn(i, 0)) {
                if (lastmask & (1u << i)) glast[d.col[i]] = gacc[i] * dps[i];
                else if (io.gsta != nullptr) io.gsta[(int64_t)b * d.sta_ncol + d.col[i] + j] = gacc[i] * dps[i];
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
static bool lean_layout_matches(const KDesc& d) {
    for (int i = 0; i < NPAR; ++i)
        if ((DM >> i) & 1)
            if (d.col[i] != lean_col<NPAR, DM, LAYOUT>(i)) return false;
    return true;
}

static bool lean_common_ok(const KDesc& d) {
    if (opt(OPT_LEAN) == 0) return false;                // 0: always K1 / K2 (A/B experiments)
    if (d.nmul != LNM || d.nvar != 3 || d.i_prcp != 0 || d.i_tmean != 1 || d.i_pet != 2) return false;
    // the forward's ring staging addresses a step as base + t * (byte stride): 32-bit strides
    if ((int64_t)d.B * d.dyn_ncol * 4 >= (1LL << 32)) return false;
    return opt(OPT_RING) != 1;                           // 1: keep the cp.async ring kernels of K1 / K2 (A/B)
}

// Which form runs where (measured on B200, `hbv` D2, ms per kernel; forward = inference / training):
//   basins          531          2,500        5,000        22,500
//   K1s ring        0.16 / 0.16  0.24 / 0.30  0.44 / 0.66  1.76 / 3.37   one-warp CTAs + cp.async ring
//   K1s regs        0.26 / 0.27  0.30 / 0.37  0.38 / 0.53  1.43 / 2.65   128-thread CTAs + register prefetch
//   warm-up ring    0.065        0.097        0.113        0.365
//   warm-up K1      0.090        0.113        0.133        0.415
//   K2s ring        0.21         0.50         1.00         4.34          (4 wide copies per step)
//   K2s regs        0.43         0.70         1.44         4.24 + memset
// -> forward: ring form up to 3 warps per scheduler (56,832 lanes; 2 when it also stores the
//    states), register form above; the
//    states-only warm-up run and the adjoint: ring form everywhere.  HBV_B200_LEAN_SMALL (lanes)
//    and HBV_B200_LEAN_BWD_RING (0/1) override for experiments.
static bool lean_small_grid(const KDesc& d, bool storing_states = false) {
    // (with the state stores of a training run the register form already wins at 2.1 warps per
    // scheduler on the hourly model: C4 K1s 12.6 vs 14.5 ms; for inference the ring wins there,
    // 7.3 vs 10.2 ms)
    long long thr = 148LL * 4 * 32 * (storing_states ? 2 : 3);
    if (opt(OPT_LEAN_SMALL) >= 0) thr = opt(OPT_LEAN_SMALL);
    return (long long)d.B * LNM <= thr;
}
static bool lean_bwd_ring(const KDesc& d) {
    return opt(OPT_LEAN_BWD_RING) != 0 && d.dyn_ncol % 2 == 0;      // (8 B copies of the parameter runs)
}

constexpr int LRD_F = 12;    // small grids: forward ring depth (steps)
constexpr int LRD_B = 8;     // small grids: adjoint ring depth

// launch with SMEM bytes of dynamic shared memory; above 48 KB the kernel is opted in once per
// process (the static is per kernel: the kernel is a template argument, not a function argument)
template <auto KERN, size_t SMEM, class D, class IO>
static void lean_go(int grid, int block, cudaStream_t st, const D& d, const IO& io) {
    if constexpr (SMEM > 48 * 1024) {
        static const cudaError_t attr = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        (void)attr;
    }
    KERN<<<grid, block, SMEM, st>>>(d, io);
}

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, int LBPB, int RD>
static int launch_fwd_lean_b(KDesc d, const FwdPtrs& io, cudaStream_t st) {
    constexpr int ND = DynSet<Traits<VAR>::NPAR, DM>::NDYN;
    d.BPB = LBPB;
    constexpr int slots = RD > 0 ? RD : (RD < 0 ? LNCH * LTC : 0);
    constexpr size_t smem = ((size_t)LTC * LBPB * (LNM * NFP + 12) + (size_t)slots * (4 * LBPB + LBPB * LNM * ND)) * sizeof(float);
    const int grid = (d.B + LBPB - 1) / LBPB;
    if constexpr (LBPB == 8 && RD < 0) {
        if (io.ckpt != nullptr && d.ck_layout == 1) {      // (try_fwd_lean let in K = 1 / 4 only)
            if (d.K == 1) lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, LBPB, RD, true, 1, 1>, smem>(grid, LBPB * LNM, st, d, io);
            else lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, LBPB, RD, true, 4, 1>, smem>(grid, LBPB * LNM, st, d, io);
            count_launch();
            count_lean_launch();
            cudaError_t e1 = cudaGetLastError();
            if (e1 != cudaSuccess) set_error(cudaGetErrorString(e1));
            return (int)e1;
        }
    }
    if (io.ckpt == nullptr) lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, false, LBPB, RD>, smem>(grid, LBPB * LNM, st, d, io);
    else if (d.K == 1) lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, LBPB, RD>, smem>(grid, LBPB * LNM, st, d, io);
    else if (d.K == 2) lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, LBPB, RD, true, 2>, smem>(grid, LBPB * LNM, st, d, io);
    else lean_go<hbv_fwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, true, LBPB, RD, true, 4>, smem>(grid, LBPB * LNM, st, d, io);
    count_launch();
    count_lean_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

// warm-up / initialize run: all parameters time-invariant, states only
template <int VAR, bool BETAET, bool SIG, int LBPB, int RD>
static int launch_fwd_lean_warm(KDesc d, const FwdPtrs& io, cudaStream_t st) {
    d.BPB = LBPB;
    const size_t smem = (size_t)RD * 8 * sizeof(float);
    const int grid = (d.B + LBPB - 1) / LBPB;
    hbv_fwd_lean_kernel<VAR, BETAET, 0, 0, SIG, false, LBPB, RD, false><<<grid, LBPB * LNM, smem, st>>>(d, io);
    count_launch();
    count_lean_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET>
int try_fwd_lean_warm(const KDesc& d, const FwdPtrs& io, cudaStream_t st) {
    if (!lean_common_ok(d)) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.state_series != nullptr || io.ckpt != nullptr)
        return HBV_NOT_ELIGIBLE;
    return d.apply_sigmoid ? launch_fwd_lean_warm<VAR, BETAET, true, 2, LRD_F>(d, io, st)
                           : launch_fwd_lean_warm<VAR, BETAET, false, 2, LRD_F>(d, io, st);
}

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG>
static int launch_fwd_lean(const KDesc& d, const FwdPtrs& io, cudaStream_t st) {
    // the warp-major state store is compiled for the chunk-ring form with K = 1 / 4 (hbv_common.cuh;
    // what the host-side policy asks for); anything else with that layout takes K1
    const bool wm = io.ckpt != nullptr && d.ck_layout != 0;
    if (wm && d.K != 1 && d.K != 4) return HBV_NOT_ELIGIBLE;
    if (lean_small_grid(d, io.ckpt != nullptr)) {
        if (wm) return HBV_NOT_ELIGIBLE;
        return launch_fwd_lean_b<VAR, BETAET, DM, LAYOUT, SIG, 2, LRD_F>(d, io, st);
    }
    // 128-thread CTAs: inputs through the chunk ring (cp.async, 8 B copies of the parameter runs)
    // up to `lean_deep` lanes, register prefetch above / when the rows are not 8 B aligned
    const long long deep_max = opt(OPT_LEAN_DEEP) >= 0 ? opt(OPT_LEAN_DEEP) : (1LL << 62);
    // (measured on B200, training forward, chunk ring / register form, ms: `hbv` D2 2,500 basins
    // 0.29 / 0.37, 4,000 0.41 / 0.46, 8,000 0.57 / 0.73, 22,500 (K = 4) 1.54 / 1.74; hourly D3 2,500
    // units x 17,520 h 8.0 / 9.1.  CTAs of 2 or 4 basins instead of 8 were measured slower on the
    // hourly grid — step 24.1 / 22.7 / 22.4 ms — and equal within noise on the `hbv` ones.)
    if ((long long)d.B * LNM <= deep_max && d.dyn_ncol % 2 == 0 && reinterpret_cast<uintptr_t>(io.dyn) % 8 == 0)
        return launch_fwd_lean_b<VAR, BETAET, DM, LAYOUT, SIG, 8, -1>(d, io, st);
    if (io.ckpt != nullptr && d.ck_layout != 0) return HBV_NOT_ELIGIBLE;
    return launch_fwd_lean_b<VAR, BETAET, DM, LAYOUT, SIG, 8, 0>(d, io, st);
}

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG, int LBPB, int PF, int RD>
static int launch_bwd_lean_b(KDesc d, const BwdPtrs& io, cudaStream_t st) {
    constexpr int ND = DynSet<Traits<VAR>::NPAR, DM>::NDYN;
    d.BPB = LBPB;
    const size_t smem = (size_t)RD * (8 + 32 * (ND + 5)) * sizeof(float);
    const int grid = (d.B + LBPB - 1) / LBPB;
    if constexpr (LBPB == 2) {
        const bool zf = io.zero_fill && popc_c((unsigned)DM) * LNM != d.dyn_ncol;
        if (d.ck_layout == 1) {          // (try_bwd_lean let in K = 1 / 4 only)
            if (d.K == 1) {
                if (zf) hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, true, 1, 1><<<grid, LBPB * LNM, smem, st>>>(d, io);
                else hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false, 1, 1><<<grid, LBPB * LNM, smem, st>>>(d, io);
            } else {
                if (zf) hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, true, 4, 1><<<grid, LBPB * LNM, smem, st>>>(d, io);
                else hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false, 4, 1><<<grid, LBPB * LNM, smem, st>>>(d, io);
            }
        } else if (d.K == 1) {
            if (zf) hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, true><<<grid, LBPB * LNM, smem, st>>>(d, io);
            else hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false><<<grid, LBPB * LNM, smem, st>>>(d, io);
        } else if (d.K == 2) {
            if (zf) hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, true, 2><<<grid, LBPB * LNM, smem, st>>>(d, io);
            else hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false, 2><<<grid, LBPB * LNM, smem, st>>>(d, io);
        } else {
            if (zf) hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, true, 4><<<grid, LBPB * LNM, smem, st>>>(d, io);
            else hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false, 4><<<grid, LBPB * LNM, smem, st>>>(d, io);
        }
    } else {
        hbv_bwd_lean_kernel<VAR, BETAET, DM, LAYOUT, SIG, LBPB, PF, RD, false><<<grid, LBPB * LNM, smem, st>>>(d, io);
    }
    count_launch();
    count_lean_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET, int DM, int LAYOUT, bool SIG>
static int launch_bwd_lean(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    const bool aligned = reinterpret_cast<uintptr_t>(io.dyn) % 8 == 0 && reinterpret_cast<uintptr_t>(io.ckpt) % 16 == 0;
    const bool ring = lean_bwd_ring(d) && aligned;
    // warp-major state store: compiled for the ring form with K = 1 / 4; anything else takes K2
    if (d.ck_layout != 0 && (!ring || (d.K != 1 && d.K != 4))) return HBV_NOT_ELIGIBLE;
    return ring ? launch_bwd_lean_b<VAR, BETAET, DM, LAYOUT, SIG, 2, 1, LRD_B>(d, io, st)
                : launch_bwd_lean_b<VAR, BETAET, DM, LAYOUT, SIG, 8, 1, 0>(d, io, st);
}

template <int VAR, bool BETAET, int DM>
int try_fwd_lean(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if (!write_flux || !lean_common_ok(d)) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.state_series != nullptr) return HBV_NOT_ELIGIBLE;
    if (io.ckpt != nullptr && d.K != 1 && d.K != 2 && d.K != 4) return HBV_NOT_ELIGIBLE;
    // the ring form copies the parameter runs 8 B at a time
    if (lean_small_grid(d, io.ckpt != nullptr) && (d.dyn_ncol % 2 != 0 || reinterpret_cast<uintptr_t>(io.dyn) % 8 != 0)) return HBV_NOT_ELIGIBLE;
    for (int f = 0; f < Traits<VAR>::NFLUX; ++f)
        if (io.flux[f] == nullptr) return HBV_NOT_ELIGIBLE;
    const bool sig = d.apply_sigmoid != 0;
    if (lean_layout_matches<NPAR, DM, 0>(d))
        return sig ? launch_fwd_lean<VAR, BETAET, DM, 0, true>(d, io, st) : launch_fwd_lean<VAR, BETAET, DM, 0, false>(d, io, st);
    if (lean_layout_matches<NPAR, DM, 1>(d))
        return sig ? launch_fwd_lean<VAR, BETAET, DM, 1, true>(d, io, st) : launch_fwd_lean<VAR, BETAET, DM, 1, false>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

template <int VAR, bool BETAET, int DM>
int try_bwd_lean(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if ((d.K != 1 && d.K != 2 && d.K != 4) || !lean_common_ok(d)) return HBV_NOT_ELIGIBLE;
    if (io.drop != nullptr || io.muwts != nullptr || io.gmuwts != nullptr || io.gforcing != nullptr ||
        io.gstate_series != nullptr || io.gdyn == nullptr) return HBV_NOT_ELIGIBLE;
    // K = 2, 4: the segment sweep exists in the ring form only
    if (d.K != 1 && !(lean_bwd_ring(d) && reinterpret_cast<uintptr_t>(io.dyn) % 8 == 0 &&
                      reinterpret_cast<uintptr_t>(io.ckpt) % 16 == 0)) return HBV_NOT_ELIGIBLE;
    // the caller asked for every element to be written (gdyn_zero_fill): nothing to do when every
    // column of `dyn` is a time-varying parameter (split form); otherwise the one-warp form zeroes
    // its rows itself (even row width: 8 B stores)
    if (io.zero_fill && popc_c((unsigned)DM) * LNM != d.dyn_ncol &&
        !(lean_bwd_ring(d) && reinterpret_cast<uintptr_t>(io.dyn) % 8 == 0 &&
          reinterpret_cast<uintptr_t>(io.ckpt) % 16 == 0)) return HBV_NOT_ELIGIBLE;
    if (io.gflux[HBV_F_QSIM] == nullptr) return HBV_NOT_ELIGIBLE;
    for (int f = 1; f < HBV_MAX_FLUX; ++f)
        if (io.gflux[f] != nullptr) return HBV_NOT_ELIGIBLE;
    const bool sig = d.apply_sigmoid != 0;
    if (lean_layout_matches<NPAR, DM, 0>(d))
        return sig ? launch_bwd_lean<VAR, BETAET, DM, 0, true>(d, io, st) : launch_bwd_lean<VAR, BETAET, DM, 0, false>(d, io, st);
    if (lean_layout_matches<NPAR, DM, 1>(d))
        return sig ? launch_bwd_lean<VAR, BETAET, DM, 1, true>(d, io, st) : launch_bwd_lean<VAR, BETAET, DM, 1, false>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

// the compiled (variant, dynamic set) pairs: the sets hbv_fwd.cu / hbv_bwd.cu specialise, minus
// all-dynamic hbv_1_1p (HBM-bound: hbv_dense.cu)
#if HBV_IN_PART(1)
template int try_fwd_lean<HBV_VARIANT_HBV, true, DM_D2>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_lean<HBV_VARIANT_HBV11P, true, DM_D2>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_lean_warm<HBV_VARIANT_HBV, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_lean_warm<HBV_VARIANT_HBV, false>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_lean_warm<HBV_VARIANT_HBV11P, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_lean_warm<HBV_VARIANT_HBV2, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
template int try_fwd_lean_warm<HBV_VARIANT_HOURLY, true>(const KDesc&, const FwdPtrs&, cudaStream_t);
#endif
#if HBV_IN_PART(2)
template int try_fwd_lean<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_lean<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
#endif
#if HBV_IN_PART(3)
template int try_bwd_lean<HBV_VARIANT_HBV, true, DM_D2>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_lean<HBV_VARIANT_HBV11P, true, DM_D2>(const KDesc&, const BwdPtrs&, cudaStream_t);
#endif
#if HBV_IN_PART(4)
template int try_bwd_lean<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_lean<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);
#endif

}  // namespace hbv
