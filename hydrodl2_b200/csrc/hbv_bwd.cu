// hbv_bwd.cu — K2: checkpointed hand-written adjoint of the HBV recurrence.
//
// Replaces PyTorch autograd over the unrolled loop (≈71*T tape nodes, CopySlices backward
// superlinear in T — SURVEY.md §3c).  The forward kernel stored the 5 states every K steps
// (hbv_fwd.cu); here each lane walks the segments last-to-first:
//   pass A  reload the checkpoint, re-run the K-1 forward steps of the segment, pushing the
//           state *before* every step onto a per-lane stack in shared memory
//           ([k][thread][state]: 5-word thread stride -> conflict-free, K*20 B per lane);
//   pass B  pop the states in reverse, re-evaluate the step with its intermediates in
//           registers and apply the adjoint (hbv_step.cuh: step_bwd).
// Gradients of time-varying (dynamic) parameters are written per step straight into the
// caller's packed gradient tensor (column i*nmul + j, through sigmoid' and the affine
// descale); time-invariant ones accumulate in registers and are written once.
#include <atomic>
#include <cstdlib>
#include "hbv_common.cuh"

namespace hbv {

// The order in which the sweep consumes time steps, known in advance: for each segment (last to
// first) pass A walks t0 .. t0+len-2 forward, pass B walks t0+len-1 .. t0 backward.  The input
// prefetch (ring or register) simply runs ahead along this sequence.
struct Sched {
    int seg, k, len, t0, K, T;
    bool passB;
    __device__ __forceinline__ void set_seg() {
        t0 = seg * K; len = min(T - t0, K); passB = (len <= 1); k = 0;
    }
    __device__ __forceinline__ void init(int nseg, int K_, int T_) { K = K_; T = T_; seg = nseg - 1; set_seg(); }
    // time index of the next request (and whether it belongs to pass B); -1 when exhausted
    __device__ __forceinline__ int next(bool& isB) {
        if (seg < 0) { isB = false; return -1; }
        int t;
        isB = passB;
        if (!passB) { t = t0 + k; if (++k == len - 1) { passB = true; k = 0; } }
        else { t = t0 + len - 1 - k; if (++k == len) { --seg; if (seg >= 0) set_seg(); } }
        return t;
    }
};

template <int VAR, bool BETAET, int DM, bool RING>
__global__ void __launch_bounds__(128, RING ? 1 : 4)   // RING: small grids, registers are free
hbv_bwd_kernel(const KDesc d, const BwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    using RS = RingSlots<NPAR, DM>;
    constexpr int GQ = RS::FIRST_FREE;            // slot of the prefetched dL/dQsim[t, b]
    constexpr int CK = GQ + 1;                    // K == 1: the 5 stored states of step t
    constexpr int NSP = (CK + 5) | 1;             // floats per thread per ring step (odd)
    extern __shared__ __align__(16) float smem[];  // state stack [K][NT][5] | input ring

    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int nmul = d.nmul;
    const int K = d.K;
    const int bl = tid / nmul;
    const int j = tid - bl * nmul;
    const int b_raw = blockIdx.x * d.BPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * nmul + j;
    const int64_t nlane = (int64_t)d.B * nmul;
    const int64_t ckb = ck_base(d, lane), ckp = ck_plane(d);      // stored-state layout (hbv_common.cuh)

    LaneConst lc;
    lc.nearzero = d.nearzero; lc.dt = d.dt; lc.inv_dt = d.inv_dt;
    lc.Ac = 0.f; lc.Elev = 0.f; lc.lfexp = 0.f;
    if constexpr (TR::LAT) init_lane_const(lc, __ldg(io.attrs + b), __ldg(io.attrs + d.B + b));

    float p[NPAR], dpd[NPAR], gacc[NPAR];
    uint32_t lastmask = 0;
    const uint32_t dynmask = resolve_params<NPAR, DM>(d, io.dyn, io.sta, io.drop, b, j, p, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < NPAR; ++i) dpd[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NPAR; ++i) gacc[i] = 0.f;
    const float* dyn_lane = io.dyn + (int64_t)b * d.dyn_ncol + j;
    const int64_t dyn_tstride = (int64_t)d.B * d.dyn_ncol;

    float gS[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) gS[s] = io.gstate_out ? __ldg(io.gstate_out + s * nlane + lane) : 0.f;

    const float* fptr = io.forcing + (int64_t)b * d.nvar;
    const int64_t f_tstride = (int64_t)d.B * d.nvar;
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;
    const float inv_nmul = 1.0f / (float)nmul;
    float* gdyn_lane = io.gdyn + (int64_t)b * d.dyn_ncol + j;
    float* const my_stack = smem + tid * 5;        // + k * NT * 5
    const int stack_step = NT * 5;

    // The usual training case — upstream gradient on the Qsim series only — gets that one value
    // per step prefetched with the inputs; any other set of series is loaded where it is used.
    bool only_q = (io.gflux[HBV_F_QSIM] != nullptr) && (mu_lane == nullptr);
#pragma unroll
    for (int f = 1; f < HBV_MAX_FLUX; ++f) only_q = only_q && (io.gflux[f] == nullptr);
    const float* gq_lane = only_q ? io.gflux[HBV_F_QSIM] + b : nullptr;

    // ---- fused zero-fill of the dense gradient tensor (the contract returns d/d(parameters)
    // with the full [T, B, ncol] shape): this CTA's BPB rows of time t are one contiguous run of
    // BPB*ncol floats; thread `tid` owns elements tid + k*NT and zeroes those that are not a
    // (non-dropped) dynamic column.  The element -> column map is resolved once here.
    uint32_t zmask = 0;
    if (io.zero_fill) {
        const int b0 = blockIdx.x * d.BPB;
        const int nz = min(d.BPB, d.B - b0) * d.dyn_ncol;
        for (int k = 0; k < 32; ++k) {
            const int idx = tid + k * NT;
            if (idx >= nz) break;
            const int row = idx / d.dyn_ncol;
            const int col = idx - row * d.dyn_ncol;
            bool isdyn = false;
            for (int i = 0; i < d.n_par; ++i)
                if (d.src[i] == HBV_SRC_DYN_T && col >= d.col[i] && col < d.col[i] + nmul)
                    isdyn = !(DM < 0 && io.drop != nullptr && io.drop[(int64_t)i * d.B + b0 + row]);
            if (!isdyn) zmask |= (1u << k);
        }
    }
    float* zrow = io.gdyn + (int64_t)blockIdx.x * d.BPB * d.dyn_ncol + tid;

    float F[HBV_MAX_FLUX];
    float gmu_acc = 0.f;
    // per-basin sum of the forcing gradient over the nmul components: shuffles when the lanes of
    // a basin are an aligned power-of-two group inside a warp; otherwise through a shared-memory
    // slab summed in component order by the basin's first lane (deterministic, like every other
    // reduction in this library — no atomics)
    const bool shfl_reduce = (nmul & (nmul - 1)) == 0 && nmul <= 32 && (NT % 32 == 0 || NT <= 32);
    float* const fx_slab = smem + d.slack;      // [NT][3] (host: d.slack = offset in floats, 0 = unused)

    // One reverse step = (1) `tape_step`: re-evaluate the forward step from its stored state and
    // keep its intermediates, (2) `adj_step`: apply the adjoint.  (1) of step t-1 does not depend
    // on (2) of step t, which the every-state-stored sweep below exploits.
    struct StepCtx {
        float p[NPAR], dpd[NPAR];
        Tape tp;
        float PET, Fq;
        float gF[HBV_MAX_FLUX];
    };
    auto tape_step = [&](const auto& in, int t, const float* st, StepCtx& c) {
        // upstream gradients of the nmul-reduced series (broadcast over the components)
#pragma unroll
        for (int f = 0; f < HBV_MAX_FLUX; ++f) c.gF[f] = 0.f;
        if (only_q) {
            c.gF[HBV_F_QSIM] = in[GQ] * inv_nmul;
        } else {
            const int64_t o = (int64_t)t * d.B + b;
#pragma unroll
            for (int f = 0; f < HBV_MAX_FLUX; ++f)
                if (f < TR::NFLUX && io.gflux[f] != nullptr) c.gF[f] = __ldg(io.gflux[f] + o) * inv_nmul;
            if (mu_lane != nullptr && io.gflux[HBV_F_QSIM] != nullptr)
                c.gF[HBV_F_QSIM] = __ldg(io.gflux[HBV_F_QSIM] + o) * __ldg(mu_lane + (int64_t)t * d.muwts_t_stride);
        }
        float S[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) S[s] = st[s];
        ring_apply_dyn<NPAR, DM>(d, dynmask, in, c.p, c.dpd);
        float P = in[0], PET = in[2];
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float Fl[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, true>(S, c.p, P, in[1], PET, lc, Fl, c.tp);
        c.PET = PET;
        c.Fq = Fl[HBV_F_QSIM];
    };
    auto adj_step = [&](const StepCtx& c, int t) {
        if (io.gstate_series != nullptr) {
            const float* gs = io.gstate_series + (int64_t)t * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) gS[s] += __ldg(gs + (int64_t)s * d.T * nlane);
        }
        float gp[NPAR];
#pragma unroll
        for (int i = 0; i < NPAR; ++i) gp[i] = 0.f;
        float gX[3];
        step_bwd<VAR, BETAET>(gS, c.gF, c.p, c.PET, lc, c.tp, gp, gX);

        if (io.gmuwts != nullptr && io.gflux[HBV_F_QSIM] != nullptr) {
            // Qsim_out = sum_j muwts_j * Qsim_j (hbv.py:511): d/d muwts_j = dL/dQsim_out * Qsim_j
            const float gm = __ldg(io.gflux[HBV_F_QSIM] + (int64_t)t * d.B + b) * c.Fq;
            if (d.muwts_t_stride != 0) { if (valid) io.gmuwts[(int64_t)t * d.muwts_t_stride + lane] = gm; }
            else gmu_acc += gm;
        }
        if (io.gforcing != nullptr) {
            if constexpr (TR::HOURLY) { gX[0] *= d.inv_dt; gX[2] *= d.inv_dt; }
            float* gx = io.gforcing + ((int64_t)t * d.B + b) * d.nvar;
            if (shfl_reduce) {          // the nmul lanes of a basin are an aligned group of a warp
                for (int o = nmul >> 1; o > 0; o >>= 1) {
                    gX[0] += __shfl_xor_sync(0xffffffffu, gX[0], o);
                    gX[1] += __shfl_xor_sync(0xffffffffu, gX[1], o);
                    gX[2] += __shfl_xor_sync(0xffffffffu, gX[2], o);
                }
                if (valid && j == 0) { gx[d.i_prcp] = gX[0]; gx[d.i_tmean] = gX[1]; gx[d.i_pet] = gX[2]; }
            } else {
                // (every thread of the CTA runs the same number of steps: the barriers are uniform)
                fx_slab[tid * 3 + 0] = gX[0]; fx_slab[tid * 3 + 1] = gX[1]; fx_slab[tid * 3 + 2] = gX[2];
                __syncthreads();
                if (valid && j == 0) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
                    for (int q = 0; q < nmul; ++q) {
                        a0 += fx_slab[(tid + q) * 3 + 0]; a1 += fx_slab[(tid + q) * 3 + 1]; a2 += fx_slab[(tid + q) * 3 + 2];
                    }
                    gx[d.i_prcp] = a0; gx[d.i_tmean] = a1; gx[d.i_pet] = a2;
                }
                __syncthreads();
            }
        }

        if (zmask != 0 && t < d.T - 1) {
            float* z = zrow + (int64_t)t * dyn_tstride;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if ((zmask >> k) == 0) break;
                if ((zmask >> k) & 1u) z[k * NT] = 0.f;
            }
        }
        float* gr = gdyn_lane + (int64_t)t * dyn_tstride;
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (DS::is_dyn(i, dynmask)) { if (valid) gr[(unsigned)d.col[i]] = gp[i] * c.dpd[i]; }
            else gacc[i] += gp[i];
        }
    };
    StepCtx cA;
#pragma unroll
    for (int i = 0; i < NPAR; ++i) { cA.p[i] = p[i]; cA.dpd[i] = 0.f; }
    auto rev_step = [&](const auto& in, int t, const float* st) {
        tape_step(in, t, st, cA);
        adj_step(cA, t);
    };
    auto fwd_only = [&](const auto& in, float (&S)[5]) {     // pass A: the same parameter copy
        ring_apply_dyn<NPAR, DM>(d, dynmask, in, cA.p, nullptr);
        float P = in[0], PET = in[2];
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        step_fwd<VAR, BETAET, false>(S, cA.p, P, in[1], PET, lc, F, cA.tp);
    };

    // ---- input prefetch along the sweep's schedule -------------------------------------------
    const int nseg = (d.T + K - 1) / K;
    Sched sch;
    sch.init(nseg, K, d.T);
    // RING: per-thread shared-memory ring, one step per cp.async group, nstage - 1 steps ahead
    const int nstage = d.nstage;
    const int step_floats = NT * NSP;
    float* const ring0 = smem + K * stack_step + tid * NSP;
    float* const ring_end = ring0 + nstage * step_floats;
    float* wp = ring0;
    const float* rp = ring0;
    auto ring_issue = [&]() {
        bool isB;
        const int t = sch.next(isB);
        if (t >= 0) {
            ring_issue_step<NPAR, DM>(d, fptr + (int64_t)t * f_tstride, dyn_lane + (int64_t)t * dyn_tstride,
                                      dynmask, wp);
            if (isB && only_q) cp_async4(wp + GQ, gq_lane + (int64_t)t * d.B);
            if (K == 1) {       // every state was stored: stage the state before step t as well
                const float* ck = io.ckpt + ckb + (int64_t)t * 5 * ckp;
#pragma unroll
                for (int s = 0; s < 5; ++s) cp_async4(wp + CK + s, ck + s * ckp);
            }
        }
        wp += step_floats;
        if (wp == ring_end) wp = ring0;
        cp_async_commit();
    };
    auto ring_pop = [&]() -> const float* {       // wait for the oldest step (ring_issue() then
                                                  // refills the slot consumed one step earlier)
        ring_wait(nstage);
        const float* cur = rp;
        rp += step_floats;
        if (rp == ring_end) rp = ring0;
        return cur;
    };
    // registers: distance-1 prefetch; `nxt` is requested one full step before it is consumed and
    // handed over by a register copy, so the (large) step body exists once
    // (K == 1: the stored state before step t travels with the inputs of step t, slots CK..CK+4)
    struct RegInG { float v[CK + 5]; __device__ __forceinline__ float operator[](int k) const { return v[k]; } };
    RegInG nxt;
    auto reg_request = [&]() {
        bool isB;
        const int t = sch.next(isB);
        if (t >= 0) {
            RegIn<NPAR, DM> tmp;
            reg_load_step<NPAR, DM>(d, fptr + (int64_t)t * f_tstride, dyn_lane + (int64_t)t * dyn_tstride,
                                    dynmask, tmp);
#pragma unroll
            for (int q = 0; q < GQ; ++q) nxt.v[q] = tmp.v[q];
            if (isB && only_q) nxt.v[GQ] = __ldg(gq_lane + (int64_t)t * d.B);
            if (K == 1) {
                const float* ck = io.ckpt + ckb + (int64_t)t * 5 * ckp;
#pragma unroll
                for (int s = 0; s < 5; ++s) nxt.v[CK + s] = __ldg(ck + s * ckp);
            }
        }
    };
    if constexpr (RING) {
        for (int s = 0; s < nstage - 1; ++s) ring_issue();
    } else {
        reg_request();
    }

    if (RING && K == 1) {
        // Every state stored (small problems): no recompute pass, the states arrive through the
        // ring with the inputs.  One warp owns a scheduler here, so the sweep is software
        // pipelined: the forward re-evaluation of step t-1 is issued alongside the adjoint of
        // step t (two independent dependency chains), with two register contexts swapping roles.
        if constexpr (RING) {
            StepCtx cB;
#pragma unroll
            for (int i = 0; i < NPAR; ++i) { cB.p[i] = p[i]; cB.dpd[i] = 0.f; }
            int t = d.T - 1;
            {
                const float* cur = ring_pop();
                ring_issue();
                tape_step(cur, t, cur + CK, cA);
            }
            for (; t >= 1; t -= 2) {
                {
                    const float* nx = ring_pop();
                    ring_issue();
                    tape_step(nx, t - 1, nx + CK, cB);
                    adj_step(cA, t);
                }
                if (t >= 2) {
                    const float* nx = ring_pop();
                    ring_issue();
                    tape_step(nx, t - 2, nx + CK, cA);
                }
                adj_step(cB, t - 1);
            }
            if (t == 0) adj_step(cA, 0);
        }
    } else if (K == 1) {
        // Every state stored, throughput regime: no recompute pass and no state stack; the
        // stored state of step t - 1 is requested with its inputs one full step ahead.
        if constexpr (!RING) {
#pragma unroll 1
            for (int t = d.T - 1; t >= 0; --t) {
                const RegInG cur = nxt;
                reg_request();
                rev_step(cur, t, &cur.v[CK]);
            }
        }
    } else
    for (int seg = nseg - 1; seg >= 0; --seg) {
        const int t0 = seg * K;
        const int len = min(d.T - t0, K);
        float S[5];
        {
            const float* ck = io.ckpt + ckb + (int64_t)seg * 5 * ckp;
#pragma unroll
            for (int s = 0; s < 5; ++s) S[s] = __ldg(ck + s * ckp);
        }
        // ---- pass A: recompute the segment, push the state before every step ---------------
        float* st = my_stack;
        for (int k = 0; k < len; ++k) {
#pragma unroll
            for (int s = 0; s < 5; ++s) st[s] = S[s];
            st += stack_step;
            if (k < len - 1) {
                if constexpr (RING) {
                    const float* cur = ring_pop();
                    ring_issue();
                    fwd_only(cur, S);
                } else {
                    const RegInG cur = nxt;
                    reg_request();
                    fwd_only(cur, S);
                }
            }
        }
        // ---- pass B: reverse sweep ---------------------------------------------------------
        const int tl = t0 + len - 1;
#pragma unroll 1
        for (int k = 0; k < len; ++k) {
            st -= stack_step;
            if constexpr (RING) {
                const float* cur = ring_pop();
                ring_issue();
                rev_step(cur, tl - k, st);
            } else {
                const RegInG cur = nxt;
                reg_request();
                rev_step(cur, tl - k, st);
            }
        }
    }
    if constexpr (RING) cp_async_wait<0>();
    // d(par)/d(raw) of the time-invariant parameters: recomputed here instead of being kept
    // live in registers through the sweep
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, io.drop, b, j, p, dpd, &lastmask);
    if (valid) {
        float* glast = gdyn_lane + (int64_t)(d.T - 1) * dyn_tstride;
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (i < d.n_par && !DS::is_dyn(i, dynmask)) {
                if (lastmask & (1u << i)) glast[d.col[i]] = gacc[i] * dpd[i];
                else if (io.gsta != nullptr) io.gsta[(int64_t)b * d.sta_ncol + d.col[i] + j] = gacc[i] * dpd[i];
            }
        }
        if (io.gstate_in != nullptr) {
#pragma unroll
            for (int s = 0; s < 5; ++s) io.gstate_in[s * nlane + lane] = gS[s];
        }
        if (io.gmuwts != nullptr && d.muwts_t_stride == 0) io.gmuwts[lane] = gmu_acc;
    }
}

template <int VAR, bool BETAET, int DM, bool RING>
static int launch_bwd_r(const KDesc& d, const BwdPtrs& io, size_t smem, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    auto k = hbv_bwd_kernel<VAR, BETAET, DM, RING>;
    cudaError_t e;
    static std::atomic<int> optin[HBV_MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev < HBV_MAX_DEVICES && optin[dev].load(std::memory_order_acquire) == 0) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        optin[dev].store(1, std::memory_order_release);
    }
    k<<<grid, NT, smem, st>>>(d, io);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

// input path by regime, as in hbv_fwd.cu (ring for small grids, registers otherwise)
template <int VAR, bool BETAET, int DM>
static int launch_bwd_dm(KDesc d, const BwdPtrs& io, cudaStream_t st) {
    if (d.K < 1 || d.K > 64) { set_error("ckpt_interval must be in [1, 64]"); return HBV_E_CKPT; }
    const int NT = d.BPB * d.nmul;
    const size_t stack = (size_t)d.K * 5 * NT * sizeof(float);
    if (stack > 100 * 1024) { set_error("ckpt_interval * nmul too large for the shared-memory state stack"); return HBV_E_CKPT; }
    // forcing gradient with nmul not a power of two: a [NT][3] slab after everything else
    const bool pow2 = (d.nmul & (d.nmul - 1)) == 0 && d.nmul <= 32 && (NT % 32 == 0 || NT <= 32);
    const size_t fx = (io.gforcing != nullptr && !pow2) ? (size_t)3 * NT * sizeof(float) : 0;
    if constexpr (DM >= 0) {
        constexpr int NSP = (RingSlots<Traits<VAR>::NPAR, DM>::FIRST_FREE + 6) | 1;
        const long long grid = (d.B + d.BPB - 1) / d.BPB;
        const long long force = opt(OPT_RING);
        const bool ring = force >= 0 ? (force == 1) : (grid * NT <= 148LL * 4 * 32 * 2);
        if (ring) {
            size_t bytes = 0;
            d.nstage = choose_nstage(NT, NSP, 1, stack + fx, grid, &bytes);
            if (d.nstage) {
                d.slack = (int)((stack + bytes) / sizeof(float));
                return launch_bwd_r<VAR, BETAET, DM, true>(d, io, stack + bytes + fx, st);
            }
        }
    }
    d.slack = (int)(stack / sizeof(float));
    return launch_bwd_r<VAR, BETAET, DM, false>(d, io, stack + fx, st);
}

// the warm-up rows in front of gdyn: only K2p zeroes them itself; every other kernel gets them
// cleared by a memset in stream order, and is told there is nothing in front
static int clear_rows_before(const KDesc& d, BwdPtrs& io, cudaStream_t st) {
    if (io.rows_before <= 0) return 0;
    const size_t n = (size_t)io.rows_before * d.B * d.dyn_ncol;
    cudaError_t e = cudaMemsetAsync(io.gdyn - n, 0, n * sizeof(float), st);
    io.rows_before = 0;
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET>
static int launch_bwd(const KDesc& d, const BwdPtrs& io_in, cudaStream_t st) {
    BwdPtrs io = io_in;
    const int dm = static_dynmask(d, io.drop != nullptr);
    if (dm == 0) { if (int rc = clear_rows_before(d, io, st)) return rc; return launch_bwd_dm<VAR, BETAET, 0>(d, io, st); }
    if constexpr (BETAET && (VAR == HBV_VARIANT_HBV || VAR == HBV_VARIANT_HBV11P)) {
        if (dm == DM_D2) {
            int rc = try_bwd_pipe<VAR, BETAET, DM_D2>(d, io, st);                    // hbv_pipe.cu
            if (rc != HBV_NOT_ELIGIBLE) return rc;
            if ((rc = clear_rows_before(d, io, st)) != 0) return rc;
            rc = try_bwd_lean<VAR, BETAET, DM_D2>(d, io, st);                        // hbv_lean.cu
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_bwd_dm<VAR, BETAET, DM_D2>(d, io, st);
        }
    }
    if (int rc0 = clear_rows_before(d, io, st)) return rc0;
    if constexpr (VAR == HBV_VARIANT_HBV11P) {
        if (dm == DM_ALL14) {
            const int rc = try_bwd_dense<VAR, BETAET, DM_ALL14>(d, io, st);          // hbv_dense.cu
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_bwd_dm<VAR, BETAET, DM_ALL14>(d, io, st);
        }
    }
    if constexpr (VAR == HBV_VARIANT_HBV2 || VAR == HBV_VARIANT_HOURLY) {
        if (dm == DM_D3) {
            int rc = try_bwd_dense<VAR, BETAET, DM_D3>(d, io, st);
            if (rc == HBV_NOT_ELIGIBLE) rc = try_bwd_pipe<VAR, BETAET, DM_D3>(d, io, st);
            if (rc == HBV_NOT_ELIGIBLE) rc = try_bwd_lean<VAR, BETAET, DM_D3>(d, io, st);
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_bwd_dm<VAR, BETAET, DM_D3>(d, io, st);
        }
    }
    return launch_bwd_dm<VAR, BETAET, -1>(d, io, st);
}

int make_kdesc(const hbv_desc_t* desc, KDesc& d);

int bwd_dispatch(const hbv_desc_t* desc, const hbv_bwd_io_t* io, cudaStream_t st) {
    KDesc d;
    int rc = make_kdesc(desc, d);
    if (rc) return rc;
    // the adjoint keeps more live registers: cap the CTA at 128 threads
    while (d.BPB > 1 && d.BPB * d.nmul > 128) d.BPB >>= 1;
    if (d.BPB * d.nmul > 128) { set_error("nmul > 128 unsupported in backward"); return HBV_E_NMUL; }
    BwdPtrs p;
    p.forcing = io->forcing; p.dyn = io->dyn; p.sta = io->sta; p.drop = io->drop;
    p.attrs = io->attrs; p.muwts = io->muwts; p.ckpt = io->ckpt;
    p.gstate_out = io->gstate_out; p.gstate_series = io->gstate_series;
    p.gdyn = io->gdyn; p.gsta = io->gsta; p.gstate_in = io->gstate_in;
    p.gforcing = io->gforcing; p.gmuwts = io->muwts ? io->gmuwts : nullptr;
    p.zero_fill = io->gdyn_zero_fill && io->gdyn != nullptr && ((d.BPB * d.dyn_ncol + d.BPB * d.nmul - 1) / (d.BPB * d.nmul) <= 32);
    if (io->gdyn_zero_fill && !p.zero_fill) { set_error("gdyn_zero_fill unsupported for this shape (ncol/nmul > 32)"); return HBV_E_SHAPE; }
    p.rows_before = (p.zero_fill && io->gdyn_rows_before > 0) ? io->gdyn_rows_before : 0;
    if (io->gdyn_rows_before > 0 && !p.zero_fill) { set_error("gdyn_rows_before needs gdyn_zero_fill"); return HBV_E_SHAPE; }
    for (int f = 0; f < HBV_MAX_FLUX; ++f) p.gflux[f] = io->gflux[f];
    switch (desc->variant) {
        case HBV_VARIANT_HBV:
            return desc->betaet ? launch_bwd<HBV_VARIANT_HBV, true>(d, p, st)
                                : launch_bwd<HBV_VARIANT_HBV, false>(d, p, st);
        case HBV_VARIANT_HBV11P: return launch_bwd<HBV_VARIANT_HBV11P, true>(d, p, st);
        case HBV_VARIANT_HBV2: return launch_bwd<HBV_VARIANT_HBV2, true>(d, p, st);
        case HBV_VARIANT_HOURLY: return launch_bwd<HBV_VARIANT_HOURLY, true>(d, p, st);
    }
    set_error("unknown variant");
    return HBV_E_VARIANT;
}

}  // namespace hbv
