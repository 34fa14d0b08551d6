// hbv_bwd.cu — K2: checkpointed hand-written adjoint of the HBV recurrence.
//
// Replaces PyTorch autograd over the unrolled loop (≈71*T tape nodes, CopySlices backward
// superlinear in T — SURVEY.md §3c).  The forward kernel stored the 5 states every K steps
// (hbv_fwd.cu); here each lane walks the segments last-to-first:
//   pass A  reload the checkpoint, re-run the K-1 forward steps of the segment, pushing the
//           state *before* every step onto a per-lane stack in shared memory
//           ([k][state][thread] -> conflict-free, K*20 B per lane);
//   pass B  pop the states in reverse, re-evaluate the step with its intermediates in
//           registers and apply the adjoint (hbv_step.cuh: step_bwd).
// Gradients of time-varying (dynamic) parameters are written per step straight into the
// caller's packed gradient tensor (column i*nmul + j, through sigmoid' and the affine
// descale); time-invariant ones accumulate in registers and are written once.
#include "hbv_common.cuh"

namespace hbv {

template <int VAR, bool BETAET, int K, int DM>
__global__ void __launch_bounds__(128, 4)
hbv_bwd_kernel(const KDesc d, const BwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    extern __shared__ __align__(16) float stack[];   // [K][5][NT]

    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int nmul = d.nmul;
    const int bl = tid / nmul;
    const int j = tid - bl * nmul;
    const int b_raw = blockIdx.x * d.BPB + bl;
    const bool valid = b_raw < d.B;
    const int b = valid ? b_raw : d.B - 1;
    const int64_t lane = (int64_t)b * nmul + j;
    const int64_t nlane = (int64_t)d.B * nmul;

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
    float* my_stack = stack + tid;

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

    Tape tp;
    float F[HBV_MAX_FLUX];
    // distance-1 input prefetch: `nxt` is requested one full step before it is consumed; the
    // hand-over is a register copy, so the (large) step body exists once — unrolling it for an
    // A/B register pair overflowed the instruction cache (ncu: no_instruction stalls)
    StepIn<DS::NS> nxt;

    auto fwd_only = [&](const StepIn<DS::NS>& in, float (&S)[5]) {
        apply_dyn<NPAR, DM>(d, dynmask, in, p, nullptr);
        float P = in.P, PET = in.PET;
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        step_fwd<VAR, BETAET, false>(S, p, P, in.T, PET, lc, F, tp);
    };
    auto rev_step = [&](const StepIn<DS::NS>& in, int t, const float* st) {
        // upstream gradients of the nmul-reduced series (broadcast over the components)
        float gF[HBV_MAX_FLUX];
        const int64_t o = (int64_t)t * d.B + b;
#pragma unroll
        for (int f = 0; f < HBV_MAX_FLUX; ++f) {
            gF[f] = 0.f;
            if (f < TR::NFLUX && io.gflux[f] != nullptr) gF[f] = __ldg(io.gflux[f] + o) * inv_nmul;
        }
        if (mu_lane != nullptr && io.gflux[HBV_F_QSIM] != nullptr)
            gF[HBV_F_QSIM] = __ldg(io.gflux[HBV_F_QSIM] + o) * __ldg(mu_lane + (int64_t)t * d.muwts_t_stride);
        if (io.gstate_series != nullptr) {
            const float* gs = io.gstate_series + (int64_t)t * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) gS[s] += __ldg(gs + (int64_t)s * d.T * nlane);
        }
        float S[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) S[s] = st[s * NT];
        apply_dyn<NPAR, DM>(d, dynmask, in, p, dpd);
        float P = in.P, PET = in.PET;
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        step_fwd<VAR, BETAET, true>(S, p, P, in.T, PET, lc, F, tp);

        float gp[NPAR];
#pragma unroll
        for (int i = 0; i < NPAR; ++i) gp[i] = 0.f;
        step_bwd<VAR, BETAET>(gS, gF, p, PET, lc, tp, gp);

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
            if (DS::is_dyn(i, dynmask)) { if (valid) gr[d.col[i]] = gp[i] * dpd[i]; }
            else gacc[i] += gp[i];
        }
    };
    auto load_in = [&](StepIn<DS::NS>& in, int t) {
        load_step<NPAR, DM>(d, fptr, f_tstride, dyn_lane, dyn_tstride, dynmask, min(max(t, 0), d.T - 1), in);
    };

    const int nseg = (d.T + K - 1) / K;
    for (int seg = nseg - 1; seg >= 0; --seg) {
        const int t0 = seg * K;
        const int len = min(d.T - t0, K);
        float S[5];
        {
            const float* ck = io.ckpt + (int64_t)seg * 5 * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) S[s] = __ldg(ck + s * nlane);
        }
        // ---- pass A: recompute the segment, push the state before every step ---------------
        if (K > 1) load_in(nxt, t0);
        for (int k = 0; k < len; ++k) {
            float* st = my_stack + k * 5 * NT;
#pragma unroll
            for (int s = 0; s < 5; ++s) st[s * NT] = S[s];
            if (k < len - 1) {
                const StepIn<DS::NS> cur = nxt;
                load_in(nxt, t0 + k + 1);
                fwd_only(cur, S);
            }
        }
        // ---- pass B: reverse sweep ---------------------------------------------------------
        const int tl = t0 + len - 1;
        load_in(nxt, tl);
#pragma unroll 1
        for (int k = 0; k < len; ++k) {
            const StepIn<DS::NS> cur = nxt;
            load_in(nxt, tl - k - 1);
            rev_step(cur, tl - k, my_stack + (len - 1 - k) * 5 * NT);
        }
    }
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
    }
}

template <int VAR, bool BETAET, int K, int DM>
static int launch_bwd_k(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    const size_t smem = (size_t)K * 5 * NT * sizeof(float);
    auto k = hbv_bwd_kernel<VAR, BETAET, K, DM>;
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k<<<grid, NT, smem, st>>>(d, io);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET, int DM>
static int launch_bwd_dm(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    switch (d.K) {
        case 1: return launch_bwd_k<VAR, BETAET, 1, DM>(d, io, st);
        case 8: return launch_bwd_k<VAR, BETAET, 8, DM>(d, io, st);
        case 16: return launch_bwd_k<VAR, BETAET, 16, DM>(d, io, st);
        case 32: return launch_bwd_k<VAR, BETAET, 32, DM>(d, io, st);
    }
    set_error("ckpt_interval must be one of 1, 8, 16, 32");
    return HBV_E_CKPT;
}

template <int VAR, bool BETAET>
static int launch_bwd(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    const int dm = static_dynmask(d, io.drop != nullptr);
    if (dm == 0) return launch_bwd_dm<VAR, BETAET, 0>(d, io, st);
    if constexpr (BETAET) {
        if (dm == DM_D2) return launch_bwd_dm<VAR, BETAET, DM_D2>(d, io, st);
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
    p.zero_fill = io->gdyn_zero_fill && io->gdyn != nullptr && ((d.BPB * d.dyn_ncol + d.BPB * d.nmul - 1) / (d.BPB * d.nmul) <= 32);
    if (io->gdyn_zero_fill && !p.zero_fill) { set_error("gdyn_zero_fill unsupported for this shape (ncol/nmul > 32)"); return HBV_E_SHAPE; }
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
