// hbv_fwd.cu — K1: fused forward HBV recurrence.
//
// Replaces the Python time loop + per-series `.mean(-1)` of
//   models/hbv/hbv.py:423-511, hbv_1_1p.py:422-524, hbv_2.py:464-585, hbv_2_hourly.py:527-683
// (≈71 ATen kernels per step in the reference) by ONE launch:
//   * one thread per (basin, component) lane, the 5 states live in registers for all T steps;
//   * parameters are read in place from the caller's packed tensor (column i*nmul + j): a
//     half-warp reads one 64 B run per parameter, no re-layout pass; sigmoid + affine descale
//     (hbv.py:201, core/calc/utils.py:24) are fused; static parameters are descaled once;
//   * forcings and dynamic parameters of step t+PF are loaded while step t computes (register
//     prefetch ring), so the dependent chain of a step never waits on HBM — this is what the
//     531-basin configurations (1-2 warps per SM, latency-bound) need;
//   * which parameters are dynamic is a template argument for the common sets (none / the
//     reference's shipped [parBETA, parBETAET] / all), a runtime mask otherwise;
//   * the nmul aggregation is a shared-memory transpose-reduce per chunk of TC time steps:
//     each lane stores its <=12 fluxes as 3x STS.128 into a padded, conflict-free tile, then
//     (t, basin, flux-quad) work items sum the nmul components and write [T, B] planes.
#include "hbv_common.cuh"

namespace hbv {

constexpr int TC = 4;        // time steps per output chunk
constexpr int NFP = 12;      // floats per lane per step in the staging tile (3 x float4)

__host__ __device__ inline int tile_bstride(int nmul) {
    // per-basin stride in floats; +12 keeps 128-bit accesses conflict-free for nmul = 16
    return nmul * NFP + 12;
}

template <int VAR, bool BETAET, bool WRITE_FLUX, int DM>
__global__ void __launch_bounds__(256)
hbv_fwd_kernel(const KDesc d, const FwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x;
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
    const int NT = blockDim.x;
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

    // ---- double-buffered input prefetch ----------------------------------------------------
    // Two register buffers of CL steps alternate (A/B): while the steps of one are computed the
    // loads of the other are in flight.  They are separate named arrays so that ptxas arms their
    // loads on different scoreboard slots — a wait on a slot waits for every load armed on it.
    constexpr int CL = DS::CL;
    constexpr int NH = TC / CL;
    static_assert(TC % (2 * CL) == 0, "output chunk must hold an even number of prefetch buffers");
    StepIn<DS::NS> bufA[CL], bufB[CL];
    auto load_buf = [&](StepIn<DS::NS> (&buf)[CL], int t) {
#pragma unroll
        for (int u = 0; u < CL; ++u)
            load_step<NPAR, DM>(d, fptr, f_tstride, dyn_lane, dyn_tstride, dynmask, min(t + u, d.T - 1), buf[u]);
    };
    load_buf(bufA, 0);

    int ck_next = (d.K > 0 && io.ckpt != nullptr) ? 0 : 0x7fffffff;
    float* ck_ptr = io.ckpt ? io.ckpt + lane : nullptr;

    Tape tp;
    auto do_step = [&](const StepIn<DS::NS>& in, int t, int tc) {
        if (t == ck_next) {
            if (valid) {
#pragma unroll
                for (int s = 0; s < 5; ++s) ck_ptr[s * nlane] = S[s];
            }
            ck_ptr += 5 * nlane;
            ck_next += d.K;
        }
        apply_dyn<NPAR, DM>(d, dynmask, in, p, nullptr);
        float P = in.P, PET = in.PET;
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float F[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, false>(S, p, P, in.T, PET, lc, F, tp);
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
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            if (h % 2 == 0) {
                load_buf(bufB, t0 + (h + 1) * CL);
#pragma unroll
                for (int u = 0; u < CL; ++u)
                    if (h * CL + u < tcn) do_step(bufA[u], t0 + h * CL + u, h * CL + u);
            } else {
                load_buf(bufA, t0 + (h + 1) * CL);
#pragma unroll
                for (int u = 0; u < CL; ++u)
                    if (h * CL + u < tcn) do_step(bufB[u], t0 + h * CL + u, h * CL + u);
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
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

template <int VAR, bool BETAET, int DM>
static int launch_fwd_dm(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    const size_t smem = write_flux ? (size_t)TC * d.BPB * tile_bstride(d.nmul) * sizeof(float) : 0;
    cudaError_t e;
    if (write_flux) {
        auto k = hbv_fwd_kernel<VAR, BETAET, true, DM>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        k<<<grid, NT, smem, st>>>(d, io);
    } else {
        hbv_fwd_kernel<VAR, BETAET, false, DM><<<grid, NT, 0, st>>>(d, io);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

// pick the compile-time dynamic-parameter set when the runtime one matches it
template <int VAR, bool BETAET>
static int launch_fwd(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    const int dm = static_dynmask(d, io.drop != nullptr);
    if (dm == 0) return launch_fwd_dm<VAR, BETAET, 0>(d, io, write_flux, st);
    if constexpr (BETAET) {
        if (dm == DM_D2) return launch_fwd_dm<VAR, BETAET, DM_D2>(d, io, write_flux, st);
    }
    (void)NPAR;
    return launch_fwd_dm<VAR, BETAET, -1>(d, io, write_flux, st);
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
