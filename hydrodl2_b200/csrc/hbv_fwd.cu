// hbv_fwd.cu — K1: fused forward HBV recurrence.
//
// Replaces the Python time loop + per-series `.mean(-1)` of
//   models/hbv/hbv.py:423-511, hbv_1_1p.py:422-524, hbv_2.py:464-585, hbv_2_hourly.py:527-683
// (≈71 ATen kernels per step in the reference) by ONE launch:
//   * one thread per (basin, component) lane, the 5 states live in registers for all T steps;
//   * parameters are read in place from the caller's packed tensor (column i*nmul + j): a
//     half-warp reads one 64 B run per parameter, no re-layout pass; sigmoid + affine descale
//     (hbv.py:201, core/calc/utils.py:24) are fused; static parameters are descaled once;
//   * forcings and dynamic parameters are staged through a per-thread shared-memory ring with
//     cp.async (hbv_common.cuh), 2-12 steps ahead of the step that consumes them: the dependent
//     chain of a step never waits on HBM, the loads hold no registers or scoreboard slots, and
//     the consuming LDS use immediate offsets (no per-step address arithmetic);
//   * which parameters are dynamic is a template argument for the common sets (none / the
//     reference's shipped [parBETA, parBETAET] / all), a runtime mask otherwise;
//   * the nmul aggregation is a shared-memory transpose-reduce per chunk of TC time steps:
//     each lane stores its <=12 fluxes as 3x STS.128 into a padded, conflict-free tile, then
//     (t, basin, flux-quad) work items sum the nmul components and write [T, B] planes.
#include <atomic>
#include <cstdlib>
#include "hbv_common.cuh"

namespace hbv {

constexpr int TC = 4;        // time steps per output chunk

template <int VAR, bool BETAET, bool WRITE_FLUX, int DM, bool RING>
__global__ void __launch_bounds__(256)
hbv_fwd_kernel(const KDesc d, const FwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    constexpr int NSP = RingCfg<NPAR, DM>::NSP;
    constexpr int RC = RingCfg<NPAR, DM>::RC;
    extern __shared__ __align__(16) float smem[];

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

    // ---- parameters: resolve sources, load + descale the time-invariant ones once ----------
    float p[NPAR];
    const uint32_t dynmask = resolve_params<NPAR, DM>(d, io.dyn, io.sta, io.drop, b, j, p, nullptr, nullptr);
    const float* dyn_lane = io.dyn + (int64_t)b * d.dyn_ncol + j;          // row t = 0
    const int64_t dyn_tstride = (int64_t)d.B * d.dyn_ncol;

    float S[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) S[s] = __ldg(io.state_in + s * nlane + lane);

    const float* fptr = io.forcing + (int64_t)b * d.nvar;
    const int64_t f_tstride = (int64_t)d.B * d.nvar;
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;

    // ---- output staging tile + this thread's reduce item -----------------------------------
    const int bstride = tile_bstride(nmul);
    float* my_slot = smem + bl * bstride + j * NFP;   // + tc * BPB * bstride
    const int tstride_s = d.BPB * bstride;
    const float inv_nmul = 1.0f / (float)nmul;
    const int items = TC * d.BPB * 3;
    // item `tid`: (tc, basin, quad)
    const int r_q = tid % 3;
    const int r_r = tid / 3;
    const int r_bl = r_r % d.BPB;
    const int r_tc = r_r / d.BPB;
    const int r_bb = blockIdx.x * d.BPB + r_bl;
    const bool r_ok = (tid < items) && (r_bb < d.B);
    const float* r_src = smem + r_tc * tstride_s + r_bl * bstride + r_q * 4;

    auto reduce_item = [&](const float* src, int q, int64_t o) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int jj = 0; jj < nmul; ++jj) {
            const float4 v = *reinterpret_cast<const float4*>(src + jj * NFP);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float s0 = (q == 0 && io.muwts != nullptr) ? 1.0f : inv_nmul;
        const int f0 = q * 4;
        if (io.flux[f0 + 0]) io.flux[f0 + 0][o] = acc.x * s0;
        if (io.flux[f0 + 1]) io.flux[f0 + 1][o] = acc.y * inv_nmul;
        if (io.flux[f0 + 2]) io.flux[f0 + 2][o] = acc.z * inv_nmul;
        if (io.flux[f0 + 3]) io.flux[f0 + 3][o] = acc.w * inv_nmul;
    };

    // ---- inputs: shared-memory ring (RING, small grids) or register double buffer ------------
    const float* frow = fptr;       // this thread's rows of the next step to stage / load
    const float* drow = dyn_lane;
    int t_issue = 0;
    // ring (hbv_common.cuh): RC * nstage steps deep
    const int nstage = d.nstage;
    const int step_floats = NT * NSP;
    float* const ring0 = smem + (WRITE_FLUX ? TC * tstride_s : 0) + tid * NSP;
    float* const ring_end = ring0 + RC * nstage * step_floats;
    float* wp = ring0;              // next ring group to fill
    auto issue_group = [&]() {      // RC steps starting at t_issue (nothing beyond T)
#pragma unroll
        for (int u = 0; u < RC; ++u) {
            if (t_issue + u < d.T) {
                ring_issue_step<NPAR, DM>(d, frow, drow, dynmask, wp + u * step_floats);
                frow += f_tstride;
                drow += dyn_tstride;
            }
        }
        t_issue += RC;
        wp += RC * step_floats;
        if (wp == ring_end) wp = ring0;
        cp_async_commit();
    };
    const float* rp = ring0;        // next ring group to consume
    // registers: two named buffers of CL steps alternate (A/B), so that ptxas arms their loads on
    // different scoreboard slots — a wait on a slot waits for every load armed on it
    constexpr int CL = DynSet<NPAR, DM>::CL;
    constexpr int NH = TC / CL;
    static_assert(TC % (2 * CL) == 0, "output chunk must hold an even number of prefetch buffers");
    RegIn<NPAR, DM> bufA[CL], bufB[CL];
    auto load_buf = [&](RegIn<NPAR, DM> (&buf)[CL]) {
#pragma unroll
        for (int u = 0; u < CL; ++u) {
            if (t_issue < d.T) {
                reg_load_step<NPAR, DM>(d, frow, drow, dynmask, buf[u]);
                frow += f_tstride;
                drow += dyn_tstride;
            }
            ++t_issue;
        }
    };
    if constexpr (RING) {
        for (int s = 0; s < nstage - 1; ++s) issue_group();
    } else {
        load_buf(bufA);
    }

    int ck_next = (d.K > 0 && io.ckpt != nullptr) ? 0 : 0x7fffffff;
    float* ck_ptr = io.ckpt ? io.ckpt + ck_base(d, lane) : nullptr;     // (layout: hbv_common.cuh)
    const int64_t ckp = ck_plane(d);

    Tape tp;
    auto do_step = [&](const auto& in, int t, int tc) {
        if (t == ck_next) {
            if (valid) {
#pragma unroll
                for (int s = 0; s < 5; ++s) ck_ptr[s * ckp] = S[s];
            }
            ck_ptr += 5 * ckp;
            ck_next += d.K;
        }
        ring_apply_dyn<NPAR, DM>(d, dynmask, in, p, nullptr);
        float P = in[0], PET = in[2];
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float F[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, false>(S, p, P, in[1], PET, lc, F, tp);
        if (io.state_series != nullptr && valid) {
            float* ss = io.state_series + (int64_t)t * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) ss[(int64_t)s * d.T * nlane] = S[s];
        }
        if constexpr (WRITE_FLUX) {
            if (mu_lane) F[HBV_F_QSIM] *= __ldg(mu_lane + (int64_t)t * d.muwts_t_stride);
            float4* slot = reinterpret_cast<float4*>(my_slot + tc * tstride_s);
            slot[0] = make_float4(F[0], F[1], F[2], F[3]);
            slot[1] = make_float4(F[4], F[5], F[6], F[7]);
            slot[2] = make_float4(F[8], F[9], F[10], TR::NFLUX > 11 ? F[11] : 0.f);
        }
    };

    for (int t0 = 0; t0 < d.T; t0 += TC) {
        const int tcn = min(TC, d.T - t0);
        if constexpr (RING) {
#pragma unroll
            for (int u = 0; u < TC; ++u) {
                if (u % RC == 0) {
                    // the group holding step t0+u has landed; refill the group consumed last
                    ring_wait(nstage);
                    issue_group();
                }
                if (u < tcn) do_step(rp + (u % RC) * step_floats, t0 + u, u);
                if ((u + 1) % RC == 0) {
                    rp += RC * step_floats;
                    if (rp == ring_end) rp = ring0;
                }
            }
        } else {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                if (h % 2 == 0) {
                    load_buf(bufB);
#pragma unroll
                    for (int u = 0; u < CL; ++u)
                        if (h * CL + u < tcn) do_step(bufA[u], t0 + h * CL + u, h * CL + u);
                } else {
                    load_buf(bufA);
#pragma unroll
                    for (int u = 0; u < CL; ++u)
                        if (h * CL + u < tcn) do_step(bufB[u], t0 + h * CL + u, h * CL + u);
                }
            }
        }
        if constexpr (WRITE_FLUX) {
            __syncthreads();
            if (r_ok && r_tc < tcn) reduce_item(r_src, r_q, (int64_t)(t0 + r_tc) * d.B + r_bb);
            for (int it = tid + NT; it < items; it += NT) {   // only when nmul < 12
                const int q = it % 3;
                const int r = it / 3;
                const int bl2 = r % d.BPB;
                const int tc = r / d.BPB;
                const int bb = blockIdx.x * d.BPB + bl2;
                if (bb < d.B && tc < tcn)
                    reduce_item(smem + tc * tstride_s + bl2 * bstride + q * 4, q, (int64_t)(t0 + tc) * d.B + bb);
            }
            __syncthreads();
        }
    }
    if constexpr (RING) cp_async_wait<0>();
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

template <int VAR, bool BETAET, bool WRITE_FLUX, int DM, bool RING>
static int launch_fwd_r(const KDesc& d, const FwdPtrs& io, size_t smem, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    auto k = hbv_fwd_kernel<VAR, BETAET, WRITE_FLUX, DM, RING>;
    cudaError_t e;
    // opt in to > 48 KB of dynamic shared memory once per kernel (and device)
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

// Input path by regime.  Small grids (a CTA or two per SM, e.g. 531 basins): the recurrence is a
// single dependent chain per warp, only prefetch DISTANCE hides HBM latency -> shared-memory
// ring.  Large grids: many resident warps hide it and the LSU/MIO pipe is the scarcer resource
// -> register double buffer.  The runtime-mask kernels (DM = -1) always use registers.
template <int VAR, bool BETAET, bool WRITE_FLUX, int DM>
static int launch_fwd_k(KDesc d, const FwdPtrs& io, cudaStream_t st) {
    using RCfg = RingCfg<Traits<VAR>::NPAR, DM>;
    size_t tile = WRITE_FLUX ? (size_t)TC * d.BPB * tile_bstride(d.nmul) * sizeof(float) : 0;
    if constexpr (DM >= 0) {
        const long long grid = (d.B + d.BPB - 1) / d.BPB;
        const long long force = opt(OPT_RING);     // 0 / 1: override (experiments)
        const bool ring = force >= 0 ? (force == 1) : (grid * d.BPB * d.nmul <= 148LL * 4 * 32 * 2);
        if (ring) {
            size_t bytes = 0;
            d.nstage = choose_nstage(d.BPB * d.nmul, RCfg::NSP, RCfg::RC, tile, grid, &bytes);
            if (d.nstage) return launch_fwd_r<VAR, BETAET, WRITE_FLUX, DM, true>(d, io, tile + bytes, st);
        }
    }
    if (tile > 100 * 1024) { set_error("nmul too large for the shared-memory output tile"); return HBV_E_NMUL; }
    return launch_fwd_r<VAR, BETAET, WRITE_FLUX, DM, false>(d, io, tile, st);
}

// pick the compile-time dynamic-parameter set when the runtime one matches it: none (warm-up /
// all-static), the reference's shipped [parBETA, parBETAET], BASELINE.json's all-14 (hbv_1_1p)
// and [parBETA, parK0, parBETAET] (hbv_2 family); any other set, and dropout, take the runtime
// mask (DM = -1)
template <int VAR, bool BETAET>
static int launch_fwd(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    const int dm = static_dynmask(d, io.drop != nullptr);
    if (!write_flux) {
        if (dm == 0) {
            int rc = try_fwd_pipe_warm<VAR, BETAET>(d, io, st);                    // hbv_pipe.cu
            if (rc == HBV_NOT_ELIGIBLE) rc = try_fwd_lean_warm<VAR, BETAET>(d, io, st);   // hbv_lean.cu
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_fwd_k<VAR, BETAET, false, 0>(d, io, st);
        }
        return launch_fwd_k<VAR, BETAET, false, -1>(d, io, st);
    }
    if (dm == 0) return launch_fwd_k<VAR, BETAET, true, 0>(d, io, st);
    if constexpr (BETAET && (VAR == HBV_VARIANT_HBV || VAR == HBV_VARIANT_HBV11P)) {
        if (dm == DM_D2) {
            int rc = try_fwd_pipe<VAR, BETAET, DM_D2>(d, io, true, st);             // hbv_pipe.cu
            if (rc == HBV_NOT_ELIGIBLE) rc = try_fwd_lean<VAR, BETAET, DM_D2>(d, io, true, st);   // hbv_lean.cu
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_fwd_k<VAR, BETAET, true, DM_D2>(d, io, st);
        }
    }
    if constexpr (VAR == HBV_VARIANT_HBV11P) {
        if (dm == DM_ALL14) {
            const int rc = try_fwd_dense<VAR, BETAET, DM_ALL14>(d, io, true, st);   // hbv_dense.cu
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_fwd_k<VAR, BETAET, true, DM_ALL14>(d, io, st);
        }
    }
    if constexpr (VAR == HBV_VARIANT_HBV2 || VAR == HBV_VARIANT_HOURLY) {
        if (dm == DM_D3) {
            int rc = try_fwd_dense<VAR, BETAET, DM_D3>(d, io, true, st);
            if (rc == HBV_NOT_ELIGIBLE) rc = try_fwd_pipe<VAR, BETAET, DM_D3>(d, io, true, st);
            if (rc == HBV_NOT_ELIGIBLE) rc = try_fwd_lean<VAR, BETAET, DM_D3>(d, io, true, st);
            return rc != HBV_NOT_ELIGIBLE ? rc : launch_fwd_k<VAR, BETAET, true, DM_D3>(d, io, st);
        }
    }
    return launch_fwd_k<VAR, BETAET, true, -1>(d, io, st);
}

int make_kdesc(const hbv_desc_t* desc, KDesc& d);  // hbv_cabi.cu

int fwd_dispatch(const hbv_desc_t* desc, const hbv_fwd_io_t* io, cudaStream_t st) {
    KDesc d;
    int rc = make_kdesc(desc, d);
    if (rc) return rc;
    FwdPtrs p;
    p.forcing = io->forcing; p.dyn = io->dyn; p.sta = io->sta; p.drop = io->drop;
    p.attrs = io->attrs; p.muwts = io->muwts; p.state_in = io->state_in;
    p.state_out = io->state_out; p.state_series = io->state_series; p.ckpt = io->ckpt;
    bool write_flux = false;
    for (int f = 0; f < HBV_MAX_FLUX; ++f) { p.flux[f] = io->flux[f]; write_flux |= (io->flux[f] != nullptr); }
    switch (desc->variant) {
        case HBV_VARIANT_HBV:
            return desc->betaet ? launch_fwd<HBV_VARIANT_HBV, true>(d, p, write_flux, st)
                                : launch_fwd<HBV_VARIANT_HBV, false>(d, p, write_flux, st);
        case HBV_VARIANT_HBV11P: return launch_fwd<HBV_VARIANT_HBV11P, true>(d, p, write_flux, st);
        case HBV_VARIANT_HBV2: return launch_fwd<HBV_VARIANT_HBV2, true>(d, p, write_flux, st);
        case HBV_VARIANT_HOURLY: return launch_fwd<HBV_VARIANT_HOURLY, true>(d, p, write_flux, st);
    }
    set_error("unknown variant");
    return HBV_E_VARIANT;
}

}  // namespace hbv
