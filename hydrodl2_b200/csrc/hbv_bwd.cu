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

template <int VAR, bool BETAET, int K>
__global__ void __launch_bounds__(128)
hbv_bwd_kernel(const KDesc d, const BwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
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
    lc.Ac = 0.f; lc.Elev = 0.f;
    if constexpr (TR::LAT) { lc.Ac = __ldg(io.attrs + b); lc.Elev = __ldg(io.attrs + d.B + b); }

    float p[NPAR], dpd[NPAR], gacc[NPAR];
    uint32_t dynmask = 0, lastmask = 0;
    const float* dyn_lane = io.dyn + (int64_t)b * d.dyn_ncol + j;
    const int64_t dyn_tstride = (int64_t)d.B * d.dyn_ncol;
    const float* dyn_last = dyn_lane + (int64_t)(d.T - 1) * dyn_tstride;
#pragma unroll
    for (int i = 0; i < NPAR; ++i) {
        p[i] = 0.f; dpd[i] = 0.f; gacc[i] = 0.f;
        if (i < d.n_par) {
            int src = d.src[i];
            if (src == HBV_SRC_DYN_T && io.drop != nullptr && io.drop[(int64_t)i * d.B + b]) src = HBV_SRC_DYN_LAST;
            if (src == HBV_SRC_DYN_T) dynmask |= (1u << i);
            else {
                float raw;
                if (src == HBV_SRC_DYN_LAST) { raw = __ldg(dyn_last + d.col[i]); lastmask |= (1u << i); }
                else raw = __ldg(io.sta + (int64_t)b * d.sta_ncol + d.col[i] + j);
                p[i] = descale(d, i, raw);
                dpd[i] = descale_grad(d, i, raw);
            }
        }
    }

    float gS[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) gS[s] = io.gstate_out ? __ldg(io.gstate_out + s * nlane + lane) : 0.f;

    const float* fptr = io.forcing + (int64_t)b * d.nvar;
    const int64_t f_tstride = (int64_t)d.B * d.nvar;
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;
    const float inv_nmul = 1.0f / (float)nmul;
    float* gdyn_lane = io.gdyn + (int64_t)b * d.dyn_ncol + j;

    auto load_inputs = [&](int t, float& P, float& Tm, float& PET, bool want_grad) {
        const float* fr = fptr + (int64_t)t * f_tstride;
        P = __ldg(fr + d.i_prcp);
        Tm = __ldg(fr + d.i_tmean);
        PET = __ldg(fr + d.i_pet);
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        if (dynmask) {
            const float* dr = dyn_lane + (int64_t)t * dyn_tstride;
#pragma unroll
            for (int i = 0; i < NPAR; ++i)
                if (dynmask & (1u << i)) {
                    const float raw = __ldg(dr + d.col[i]);
                    p[i] = descale(d, i, raw);
                    if (want_grad) dpd[i] = descale_grad(d, i, raw);
                }
        }
    };

    Tape tp;
    float F[HBV_MAX_FLUX];
    const int nseg = (d.T + K - 1) / K;
    for (int seg = nseg - 1; seg >= 0; --seg) {
        const int t0 = seg * K;
        const int t1 = min(d.T, t0 + K);
        float S[5];
        {
            const float* ck = io.ckpt + (int64_t)seg * 5 * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) S[s] = __ldg(ck + s * nlane);
        }
        // pass A: recompute the segment, push pre-step states
        for (int t = t0; t < t1; ++t) {
            float* st = stack + (t - t0) * 5 * NT + tid;
#pragma unroll
            for (int s = 0; s < 5; ++s) st[s * NT] = S[s];
            if (t < t1 - 1) {
                float P, Tm, PET;
                load_inputs(t, P, Tm, PET, false);
                step_fwd<VAR, BETAET, false>(S, p, P, Tm, PET, lc, F, tp);
            }
        }
        // pass B: reverse sweep
        for (int t = t1 - 1; t >= t0; --t) {
            const float* st = stack + (t - t0) * 5 * NT + tid;
#pragma unroll
            for (int s = 0; s < 5; ++s) S[s] = st[s * NT];
            float P, Tm, PET;
            load_inputs(t, P, Tm, PET, true);
            step_fwd<VAR, BETAET, true>(S, p, P, Tm, PET, lc, F, tp);

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
            float gp[NPAR];
#pragma unroll
            for (int i = 0; i < NPAR; ++i) gp[i] = 0.f;
            step_bwd<VAR, BETAET>(gS, gF, p, PET, lc, tp, gp);

            float* gr = gdyn_lane + (int64_t)t * dyn_tstride;
#pragma unroll
            for (int i = 0; i < NPAR; ++i) {
                if (dynmask & (1u << i)) { if (valid) gr[d.col[i]] = gp[i] * dpd[i]; }
                else gacc[i] += gp[i];
            }
        }
    }
    if (valid) {
        float* glast = gdyn_lane + (int64_t)(d.T - 1) * dyn_tstride;
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (i < d.n_par && !(dynmask & (1u << i))) {
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

template <int VAR, bool BETAET, int K>
static int launch_bwd_k(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    const size_t smem = (size_t)K * 5 * NT * sizeof(float);
    auto k = hbv_bwd_kernel<VAR, BETAET, K>;
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

template <int VAR, bool BETAET>
static int launch_bwd(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    switch (d.K) {
        case 1: return launch_bwd_k<VAR, BETAET, 1>(d, io, st);
        case 8: return launch_bwd_k<VAR, BETAET, 8>(d, io, st);
        case 16: return launch_bwd_k<VAR, BETAET, 16>(d, io, st);
        case 32: return launch_bwd_k<VAR, BETAET, 32>(d, io, st);
    }
    set_error("ckpt_interval must be one of 1, 8, 16, 32");
    return HBV_E_CKPT;
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
