// hbv_fwd.cu — K1: fused forward HBV recurrence.
//
// Replaces the Python time loop + per-series `.mean(-1)` of
//   models/hbv/hbv.py:423-511, hbv_1_1p.py:422-524, hbv_2.py:464-585, hbv_2_hourly.py:527-683
// (≈71 ATen kernels per step in the reference) by ONE launch:
//   * one thread per (basin, component) lane, the 5 states live in registers for all T steps;
//   * parameters are read in place from the caller's packed tensor (column i*nmul + j), a
//     half-warp reads one 64 B run per parameter -> full sectors, no re-layout pass;
//     sigmoid + affine descale (hbv.py:201, core/calc/utils.py:24) are fused;
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

template <int VAR, bool BETAET, bool WRITE_FLUX>
__global__ void __launch_bounds__(256)
hbv_fwd_kernel(const KDesc d, const FwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
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
    lc.Ac = 0.f; lc.Elev = 0.f;
    if constexpr (TR::LAT) { lc.Ac = __ldg(io.attrs + b); lc.Elev = __ldg(io.attrs + d.B + b); }

    // ---- parameters: resolve sources, load the time-invariant ones once -----------------
    float p[NPAR];
    uint32_t dynmask = 0;
    const float* dyn_lane = io.dyn + (int64_t)b * d.dyn_ncol + j;          // row t = 0
    const int64_t dyn_tstride = (int64_t)d.B * d.dyn_ncol;
    const float* dyn_last = dyn_lane + (int64_t)(d.T - 1) * dyn_tstride;   // row T-1
#pragma unroll
    for (int i = 0; i < NPAR; ++i) {
        p[i] = 0.f;
        if (i < d.n_par) {
            int src = d.src[i];
            if (src == HBV_SRC_DYN_T && io.drop != nullptr && io.drop[(int64_t)i * d.B + b]) src = HBV_SRC_DYN_LAST;
            if (src == HBV_SRC_DYN_T) dynmask |= (1u << i);
            else if (src == HBV_SRC_DYN_LAST) p[i] = descale(d, i, __ldg(dyn_last + d.col[i]));
            else p[i] = descale(d, i, __ldg(io.sta + (int64_t)b * d.sta_ncol + d.col[i] + j));
        }
    }

    float S[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) S[s] = __ldg(io.state_in + s * nlane + lane);

    const float* fptr = io.forcing + (int64_t)b * d.nvar;
    const int64_t f_tstride = (int64_t)d.B * d.nvar;
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;

    const int bstride = tile_bstride(nmul);
    float* my_slot = smem + bl * bstride + j * NFP;   // + tc * BPB * bstride
    const int tstride_s = d.BPB * bstride;
    const float inv_nmul = 1.0f / (float)nmul;
    const int NT = blockDim.x;

    Tape tp;
    for (int t0 = 0; t0 < d.T; t0 += TC) {
        const int tcn = min(TC, d.T - t0);
        for (int tc = 0; tc < tcn; ++tc) {
            const int t = t0 + tc;
            if (d.K > 0 && io.ckpt != nullptr && (t % d.K) == 0 && valid) {
                float* ck = io.ckpt + (int64_t)(t / d.K) * 5 * nlane + lane;
#pragma unroll
                for (int s = 0; s < 5; ++s) ck[s * nlane] = S[s];
            }
            const float* fr = fptr + (int64_t)t * f_tstride;
            float P = __ldg(fr + d.i_prcp);
            const float Tm = __ldg(fr + d.i_tmean);
            float PET = __ldg(fr + d.i_pet);
            if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
            if (dynmask) {
                const float* dr = dyn_lane + (int64_t)t * dyn_tstride;
#pragma unroll
                for (int i = 0; i < NPAR; ++i)
                    if (dynmask & (1u << i)) p[i] = descale(d, i, __ldg(dr + d.col[i]));
            }
            float F[HBV_MAX_FLUX];
            step_fwd<VAR, BETAET, false>(S, p, P, Tm, PET, lc, F, tp);

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
        }
        if constexpr (WRITE_FLUX) {
            __syncthreads();
            const int items = tcn * d.BPB * 3;
            for (int it = tid; it < items; it += NT) {
                const int q = it % 3;
                const int r = it / 3;
                const int bl2 = r % d.BPB;
                const int tc = r / d.BPB;
                const int bb = blockIdx.x * d.BPB + bl2;
                if (bb >= d.B) continue;
                const float* src = smem + tc * tstride_s + bl2 * bstride + q * 4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int jj = 0; jj < nmul; ++jj) {
                    const float4 v = *reinterpret_cast<const float4*>(src + jj * NFP);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const int64_t o = (int64_t)(t0 + tc) * d.B + bb;
                const float s0 = (q == 0 && io.muwts != nullptr) ? 1.0f : inv_nmul;
                const int f0 = q * 4;
                if (io.flux[f0 + 0]) io.flux[f0 + 0][o] = acc.x * s0;
                if (io.flux[f0 + 1]) io.flux[f0 + 1][o] = acc.y * inv_nmul;
                if (io.flux[f0 + 2]) io.flux[f0 + 2][o] = acc.z * inv_nmul;
                if (io.flux[f0 + 3]) io.flux[f0 + 3][o] = acc.w * inv_nmul;
            }
            __syncthreads();
        }
    }
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

template <int VAR, bool BETAET>
static int launch_fwd(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    const int NT = d.BPB * d.nmul;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    const size_t smem = write_flux ? (size_t)TC * d.BPB * tile_bstride(d.nmul) * sizeof(float) : 0;
    cudaError_t e;
    if (write_flux) {
        auto k = hbv_fwd_kernel<VAR, BETAET, true>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        k<<<grid, NT, smem, st>>>(d, io);
    } else {
        hbv_fwd_kernel<VAR, BETAET, false><<<grid, NT, 0, st>>>(d, io);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
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
