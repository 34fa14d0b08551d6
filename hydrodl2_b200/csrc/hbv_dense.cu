// hbv_dense.cu — K1d / K2d: TMA-staged forms of K1 (hbv_fwd.cu) and K2 (hbv_bwd.cu) for runs
// whose time-varying parameters fill (most of) the parameter tensor's rows — BASELINE.json
// configs[2] (hbv_1_1p, all 14 parameters dynamic: 904 of the 972 B per basin-step are parameter
// bytes) and the split-form `hbv_2` family (every column of the dynamic tensor is read).
//
// In that regime a CTA's inputs of one time step are ONE contiguous run of the caller's tensor
// (BPB basins x ncol floats), so instead of 14 + 3 four-byte loads per thread-step:
//   * one thread issues `cp.async.bulk` (TMA, SASS UBLKCP) copies of the parameter run, the
//     forcing run (and, in the adjoint, the upstream-gradient run and the five stored-state
//     runs) into a shared-memory ring, DTC steps per mbarrier, 2-3 chunks ahead; the copies
//     complete on the mbarrier's transaction count — no registers, no LSU address arithmetic,
//     exact bytes from L2 (no 32 B-sector over-fetch of misaligned 64 B runs);
//   * lanes read their values back with immediate-offset LDS;
//   * the adjoint stages each step's parameter-gradient rows in shared memory and writes them
//     with ONE bulk store per step: full 16 B-aligned runs instead of 14 half-warp 64 B runs,
//     zeros for the non-parameter columns included, so the dense [T, B, ncol] gradient tensor
//     needs no memset and no read-modify-write of partial sectors.
// A run's first byte is only 4 B-aligned in general (row width 226 floats): the copy starts at
// the enclosing 16 B boundary and lanes add the 0-12 B shift when they read.  The step
// arithmetic is the same hbv_step.cuh code as K1/K2 — results are bit-identical to them.
//
// The adjoint here is the every-state-stored sweep (K = 1, hbv_b200_auto_ckpt): the state before
// step t arrives with the inputs of step t, no recompute pass.
//
// Reference spans replaced: models/hbv/hbv_1_1p.py:422-524, hbv_2.py:464-585,
// hbv_2_hourly.py:527-683 (forward) and PyTorch autograd over them (backward).
#include <atomic>
#include <cstdlib>
#include "hbv_common.cuh"

namespace hbv {

constexpr int DTC = 2;       // time steps per mbarrier / output chunk (forward)
constexpr int DNM = 16;      // components per basin this path is compiled for

__host__ __device__ inline int up16(int x) { return (x + 15) & ~15; }

// shared-memory geometry, identical on host and device
struct DenseGeom {
    int pbytes, fbytes;      // parameter / forcing run of one step (+ d.slack for the 0-12 B shift)
    int gbytes, sbytes;      // adjoint: upstream-gradient run, five stored-state runs
    int tile_bytes;          // forward: output staging tile
    int obytes;              // adjoint: one gradient staging buffer
};
__host__ __device__ inline DenseGeom dense_geom(const KDesc& d) {
    DenseGeom g;
    g.pbytes = up16(d.BPB * d.dyn_ncol * 4) + d.slack;
    g.fbytes = up16(d.BPB * d.nvar * 4) + d.slack;
    g.gbytes = up16(d.BPB * 4) + d.slack;
    g.sbytes = 5 * d.BPB * DNM * 4;
    g.tile_bytes = up16(DTC * d.BPB * tile_bstride(DNM) * 4);
    g.obytes = up16(d.BPB * d.dyn_ncol * 4) + d.slack;
    return g;
}

// column of parameter i inside a row of `dyn`: LAYOUT 0 = packed form (hbv.py:201-208: parameter
// i at i*nmul), LAYOUT 1 = split form (hbv_2.py:211-230: dynamic parameters only, in order)
template <int NPAR, int DM, int LAYOUT>
__host__ __device__ constexpr int dense_col(int i) {
    return (LAYOUT == 0 ? i : DynSet<NPAR, DM>::slot(i)) * DNM;
}

// ---- PTX: mbarrier + bulk async copies (TMA, non-tensor form) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBW_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBD_%=;\n"
        "bra MBW_%=;\n"
        "MBD_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`; dst, src 16 B aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make this thread's shared-memory writes visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one aligned global -> shared copy of the 4 B-aligned run [src, src + bytes): returns the bytes
// the mbarrier must expect; with `issue` false only computes that size
__device__ __forceinline__ uint32_t stage_run(void* dst, const char* src, uint32_t bytes, uint64_t* bar, bool issue) {
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t sz = (sh + bytes + 15u) & ~15u;
    if (issue) bulk_g2s(dst, src - sh, sz, bar);
    return sz;
}

// ================================================================================================
// K1d: forward
// ================================================================================================
template <int VAR, bool BETAET, int DM, int LAYOUT>
__global__ void __launch_bounds__(128, 4)
hbv_fwd_dense_kernel(const KDesc d, const FwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);          // <= 8 mbarriers

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b0 = blockIdx.x * d.BPB;
    const int nb = min(d.BPB, d.B - b0);
    const bool valid = bl < nb;
    const int ble = valid ? bl : nb - 1;          // lanes past the last basin shadow it (never stored)
    const int b = b0 + ble;
    const int64_t lane = (int64_t)b * DNM + j;
    const int64_t nlane = (int64_t)d.B * DNM;
    const DenseGeom g = dense_geom(d);
    float* const tile = reinterpret_cast<float*>(sm + 64);
    unsigned char* const ring = sm + 64 + g.tile_bytes;
    const int slot = g.pbytes + g.fbytes;
    const int NG = d.nstage / DTC;                // mbarriers = chunks in flight

    LaneConst lc;
    lc.nearzero = d.nearzero; lc.dt = d.dt; lc.inv_dt = d.inv_dt;
    lc.Ac = 0.f; lc.Elev = 0.f; lc.lfexp = 0.f;
    if constexpr (TR::LAT) init_lane_const(lc, __ldg(io.attrs + b), __ldg(io.attrs + d.B + b));

    float p[NPAR];
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, nullptr, nullptr);
    float S[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) S[s] = __ldg(io.state_in + s * nlane + lane);
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;

    // ---- producer side (thread 0): the CTA's runs of one time step ---------------------------
    const char* const prun = reinterpret_cast<const char*>(io.dyn + (int64_t)b0 * d.dyn_ncol);
    const char* const frun = reinterpret_cast<const char*>(io.forcing + (int64_t)b0 * d.nvar);
    const int64_t p_tstride = (int64_t)d.B * d.dyn_ncol * 4;
    const int64_t f_tstride = (int64_t)d.B * d.nvar * 4;
    const uint32_t pn = (uint32_t)nb * d.dyn_ncol * 4, fn = (uint32_t)nb * d.nvar * 4;
    auto issue_chunk = [&](int c, int grp) {      // steps c*DTC .. c*DTC+DTC-1 into group grp
        const int t_lo = c * DTC;
        if (t_lo >= d.T) return;
        uint32_t tot = 0;
#pragma unroll
        for (int u = 0; u < DTC; ++u)
            if (t_lo + u < d.T) {
                tot += stage_run(nullptr, prun + (t_lo + u) * p_tstride, pn, nullptr, false);
                tot += stage_run(nullptr, frun + (t_lo + u) * f_tstride, fn, nullptr, false);
            }
        mbar_arrive_expect_tx(&bars[grp], tot);
#pragma unroll
        for (int u = 0; u < DTC; ++u)
            if (t_lo + u < d.T) {
                unsigned char* dst = ring + (grp * DTC + u) * slot;
                stage_run(dst, prun + (t_lo + u) * p_tstride, pn, &bars[grp], true);
                stage_run(dst + g.pbytes, frun + (t_lo + u) * f_tstride, fn, &bars[grp], true);
            }
    };
    if (tid == 0) {
        for (int q = 0; q < NG; ++q) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int q = 0; q < NG; ++q) issue_chunk(q, q);

    // ---- consumer side: this lane's offsets into a ring slot ---------------------------------
    uint32_t shp = (uint32_t)(reinterpret_cast<uintptr_t>(prun) & 15u);
    uint32_t shf = (uint32_t)(reinterpret_cast<uintptr_t>(frun) & 15u);
    const uint32_t incp = (uint32_t)(p_tstride & 15), incf = (uint32_t)(f_tstride & 15);
    const unsigned char* const lane_p = ring + (ble * d.dyn_ncol + j) * 4;
    const unsigned char* const lane_f = ring + g.pbytes + ble * d.nvar * 4;

    // ---- output staging tile + this thread's reduce item (as in hbv_fwd.cu) ------------------
    const int bstride = tile_bstride(DNM);
    float* my_slot = tile + bl * bstride + j * NFP;
    const int tstride_s = d.BPB * bstride;
    constexpr float inv_nmul = 1.0f / (float)DNM;
    const int items = DTC * d.BPB * 3;
    const int r_q = tid % 3;
    const int r_r = tid / 3;
    const int r_bl = r_r % d.BPB;
    const int r_tc = r_r / d.BPB;
    const int r_bb = b0 + r_bl;
    const bool r_ok = (tid < items) && (r_bb < d.B);
    const float* r_src = tile + r_tc * tstride_s + r_bl * bstride + r_q * 4;
    auto reduce_item = [&](const float* src, int q, int64_t o) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int jj = 0; jj < DNM; ++jj) {
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

    int ck_next = (d.K > 0 && io.ckpt != nullptr) ? 0 : 0x7fffffff;
    float* ck_ptr = io.ckpt ? io.ckpt + lane : nullptr;
    Tape tp;

    const int nchunk = (d.T + DTC - 1) / DTC;
    int grp = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunk; ++c) {
        mbar_wait(&bars[grp], phase);
        const int t0 = c * DTC;
        const int tcn = min(DTC, d.T - t0);
#pragma unroll
        for (int u = 0; u < DTC; ++u) {
            if (u < tcn) {
                const int t = t0 + u;
                const float* lp = reinterpret_cast<const float*>(lane_p + (grp * DTC + u) * slot + shp);
                const float* lf = reinterpret_cast<const float*>(lane_f + (grp * DTC + u) * slot + shf);
                if (t == ck_next) {
                    if (valid) {
#pragma unroll
                        for (int s = 0; s < 5; ++s) ck_ptr[s * nlane] = S[s];
                    }
                    ck_ptr += 5 * nlane;
                    ck_next += d.K;
                }
#pragma unroll
                for (int i = 0; i < NPAR; ++i)
                    if (DS::is_dyn(i, 0)) p[i] = descale(d, i, lp[dense_col<NPAR, DM, LAYOUT>(i)]);
                float P = lf[d.i_prcp], PET = lf[d.i_pet];
                const float Tm = lf[d.i_tmean];
                if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
                float F[HBV_MAX_FLUX];
                step_fwd<VAR, BETAET, false>(S, p, P, Tm, PET, lc, F, tp);
                if (io.state_series != nullptr && valid) {
                    float* ss = io.state_series + (int64_t)t * nlane + lane;
#pragma unroll
                    for (int s = 0; s < 5; ++s) ss[(int64_t)s * d.T * nlane] = S[s];
                }
                if (mu_lane) F[HBV_F_QSIM] *= __ldg(mu_lane + (int64_t)t * d.muwts_t_stride);
                float4* o4 = reinterpret_cast<float4*>(my_slot + u * tstride_s);
                o4[0] = make_float4(F[0], F[1], F[2], F[3]);
                o4[1] = make_float4(F[4], F[5], F[6], F[7]);
                o4[2] = make_float4(F[8], F[9], F[10], TR::NFLUX > 11 ? F[11] : 0.f);
                shp = (shp + incp) & 15u;
                shf = (shf + incf) & 15u;
            }
        }
        __syncthreads();                      // tile complete; every lane is done with this group's slots
        if (tid == 0) issue_chunk(c + NG, grp);
        if (r_ok && r_tc < tcn) reduce_item(r_src, r_q, (int64_t)(t0 + r_tc) * d.B + r_bb);
        __syncthreads();                      // tile free
        if (++grp == NG) { grp = 0; phase ^= 1u; }
    }
    if (valid && io.state_out != nullptr) {
#pragma unroll
        for (int s = 0; s < 5; ++s) io.state_out[s * nlane + lane] = S[s];
    }
}

// ================================================================================================
// K2d: adjoint (every state stored, K = 1)
// ================================================================================================
// MINB: resident CTAs per SM the register allocation is sized for (4: <= 128 registers, 5: <= 96)
template <int VAR, bool BETAET, int DM, int LAYOUT, int MINB>
__global__ void __launch_bounds__(128, MINB)
hbv_bwd_dense_kernel(const KDesc d, const BwdPtrs io) {
    using TR = Traits<VAR>;
    constexpr int NPAR = TR::NPAR;
    using DS = DynSet<NPAR, DM>;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);

    const int tid = threadIdx.x;
    const int bl = tid >> 4, j = tid & 15;
    const int b0 = blockIdx.x * d.BPB;
    const int nb = min(d.BPB, d.B - b0);
    const bool valid = bl < nb;
    const int ble = valid ? bl : nb - 1;
    const int b = b0 + ble;
    const int64_t lane = (int64_t)b * DNM + j;
    const int64_t nlane = (int64_t)d.B * DNM;
    const DenseGeom g = dense_geom(d);
    unsigned char* const obuf = sm + 64;                       // 2 gradient staging buffers
    unsigned char* const ring = sm + 64 + 2 * g.obytes;
    const int slot = g.pbytes + g.fbytes + g.gbytes + g.sbytes;
    const int NS = d.nstage;

    LaneConst lc;
    lc.nearzero = d.nearzero; lc.dt = d.dt; lc.inv_dt = d.inv_dt;
    lc.Ac = 0.f; lc.Elev = 0.f; lc.lfexp = 0.f;
    if constexpr (TR::LAT) init_lane_const(lc, __ldg(io.attrs + b), __ldg(io.attrs + d.B + b));

    float p[NPAR], dpd[NPAR], gacc[NPAR];
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < NPAR; ++i) { dpd[i] = 0.f; gacc[i] = 0.f; }
    float gS[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) gS[s] = io.gstate_out ? __ldg(io.gstate_out + s * nlane + lane) : 0.f;
    const float* mu_lane = io.muwts ? io.muwts + lane : nullptr;
    constexpr float inv_nmul = 1.0f / (float)DNM;

    bool only_q = (io.gflux[HBV_F_QSIM] != nullptr) && (mu_lane == nullptr);
#pragma unroll
    for (int f = 1; f < HBV_MAX_FLUX; ++f) only_q = only_q && (io.gflux[f] == nullptr);

    // ---- producer side (thread 0) ------------------------------------------------------------
    const char* const prun = reinterpret_cast<const char*>(io.dyn + (int64_t)b0 * d.dyn_ncol);
    const char* const frun = reinterpret_cast<const char*>(io.forcing + (int64_t)b0 * d.nvar);
    const char* const qrun = only_q ? reinterpret_cast<const char*>(io.gflux[HBV_F_QSIM] + b0) : nullptr;
    const char* const srun = reinterpret_cast<const char*>(io.ckpt + (int64_t)b0 * DNM);
    const int64_t p_tstride = (int64_t)d.B * d.dyn_ncol * 4;
    const int64_t f_tstride = (int64_t)d.B * d.nvar * 4;
    const int64_t q_tstride = (int64_t)d.B * 4;
    const int64_t s_sstride = nlane * 4;                       // between the 5 states of a step
    const uint32_t pn = (uint32_t)nb * d.dyn_ncol * 4, fn = (uint32_t)nb * d.nvar * 4;
    const uint32_t qn = (uint32_t)nb * 4, sn = (uint32_t)nb * DNM * 4;
    auto issue_step = [&](int t, int sl) {                     // inputs of step t into slot sl
        if (t < 0) return;
        unsigned char* dst = ring + sl * slot;
        uint32_t tot = stage_run(nullptr, prun + t * p_tstride, pn, nullptr, false)
                     + stage_run(nullptr, frun + t * f_tstride, fn, nullptr, false) + 5 * sn;
        if (only_q) tot += stage_run(nullptr, qrun + t * q_tstride, qn, nullptr, false);
        mbar_arrive_expect_tx(&bars[sl], tot);
        stage_run(dst, prun + t * p_tstride, pn, &bars[sl], true);
        stage_run(dst + g.pbytes, frun + t * f_tstride, fn, &bars[sl], true);
        if (only_q) stage_run(dst + g.pbytes + g.fbytes, qrun + t * q_tstride, qn, &bars[sl], true);
        unsigned char* sd = dst + g.pbytes + g.fbytes + g.gbytes;
#pragma unroll
        for (int s = 0; s < 5; ++s)
            bulk_g2s(sd + s * d.BPB * DNM * 4, srun + ((int64_t)t * 5 + s) * s_sstride, sn, &bars[sl]);
    };
    if (tid == 0) {
        for (int q = 0; q < NS; ++q) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    // the staging buffers start as zeros: the columns that carry no gradient are never written
    // again, so every bulk store writes their zeros along with the gradients
    for (int k = tid * 4; k < 2 * g.obytes; k += blockDim.x * 4) *reinterpret_cast<float*>(obuf + k) = 0.f;
    __syncthreads();
    if (tid == 0)
        for (int q = 0; q < NS; ++q) issue_step(d.T - 1 - q, q);

    // ---- consumer side -----------------------------------------------------------------------
    uint32_t shp = (uint32_t)(reinterpret_cast<uintptr_t>(prun + (d.T - 1) * p_tstride) & 15u);
    uint32_t shf = (uint32_t)(reinterpret_cast<uintptr_t>(frun + (d.T - 1) * f_tstride) & 15u);
    uint32_t shq = only_q ? (uint32_t)(reinterpret_cast<uintptr_t>(qrun + (d.T - 1) * q_tstride) & 15u) : 0u;
    const uint32_t incp = (uint32_t)(p_tstride & 15), incf = (uint32_t)(f_tstride & 15), incq = (uint32_t)(q_tstride & 15);
    const unsigned char* const lane_p = ring + (ble * d.dyn_ncol + j) * 4;
    const unsigned char* const lane_f = ring + g.pbytes + ble * d.nvar * 4;
    const unsigned char* const lane_q = ring + g.pbytes + g.fbytes + ble * 4;
    const unsigned char* const lane_s = ring + g.pbytes + g.fbytes + g.gbytes + (ble * DNM + j) * 4;

    // gradient rows of this CTA: the run [grun + t * p_tstride, + pn) of the caller's tensor.  The
    // dispatcher guarantees a time-invariant 16 B phase (p_tstride % 16 == 0), so the aligned core
    // and the <= 12 B head / tail of the run sit at fixed offsets of the staging buffer.
    char* const grun = reinterpret_cast<char*>(io.gdyn + (int64_t)b0 * d.dyn_ncol);
    const uint32_t shg = (uint32_t)(reinterpret_cast<uintptr_t>(grun) & 15u);
    const uint32_t head = (16u - shg) & 15u;                   // bytes before the first 16 B boundary
    const uint32_t core = (pn > head) ? ((pn - head) & ~15u) : 0u;
    const uint32_t tail = pn - min(pn, head) - core;
    unsigned char* const lane_o = obuf + shg + (ble * d.dyn_ncol + j) * 4;
    float* const gdyn_lane = io.gdyn + (int64_t)b * d.dyn_ncol + j;

    const bool shfl_reduce = true;   // nmul = 16: the lanes of a basin are an aligned half warp
    float gmu_acc = 0.f;
    Tape tp;

    int sl = 0;
    uint32_t phase = 0;
    for (int t = d.T - 1; t >= 0; --t) {
        mbar_wait(&bars[sl], phase);
        const float* lp = reinterpret_cast<const float*>(lane_p + sl * slot + shp);
        const float* lf = reinterpret_cast<const float*>(lane_f + sl * slot + shf);
        const float* ls = reinterpret_cast<const float*>(lane_s + sl * slot);

        // upstream gradients of the nmul-reduced series (broadcast over the components)
        float gF[HBV_MAX_FLUX];
#pragma unroll
        for (int f = 0; f < HBV_MAX_FLUX; ++f) gF[f] = 0.f;
        if (only_q) {
            gF[HBV_F_QSIM] = *reinterpret_cast<const float*>(lane_q + sl * slot + shq) * inv_nmul;
        } else {
            const int64_t o = (int64_t)t * d.B + b;
#pragma unroll
            for (int f = 0; f < HBV_MAX_FLUX; ++f)
                if (f < TR::NFLUX && io.gflux[f] != nullptr) gF[f] = __ldg(io.gflux[f] + o) * inv_nmul;
            if (mu_lane != nullptr && io.gflux[HBV_F_QSIM] != nullptr)
                gF[HBV_F_QSIM] = __ldg(io.gflux[HBV_F_QSIM] + o) * __ldg(mu_lane + (int64_t)t * d.muwts_t_stride);
        }
        // re-evaluate the step from its stored state, keep the intermediates
        float S[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) S[s] = ls[s * d.BPB * DNM];
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, 0)) descale_both(d, i, lp[dense_col<NPAR, DM, LAYOUT>(i)], p[i], dpd[i]);
        float P = lf[d.i_prcp], PET = lf[d.i_pet];
        const float Tm = lf[d.i_tmean];
        if constexpr (TR::HOURLY) { P = P * d.inv_dt; PET = PET * d.inv_dt; }
        float Fl[HBV_MAX_FLUX];
        step_fwd<VAR, BETAET, true>(S, p, P, Tm, PET, lc, Fl, tp);

        if (io.gstate_series != nullptr) {
            const float* gs = io.gstate_series + (int64_t)t * nlane + lane;
#pragma unroll
            for (int s = 0; s < 5; ++s) gS[s] += __ldg(gs + (int64_t)s * d.T * nlane);
        }
        float gp[NPAR];
#pragma unroll
        for (int i = 0; i < NPAR; ++i) gp[i] = 0.f;
        float gX[3];
        step_bwd<VAR, BETAET>(gS, gF, p, PET, lc, tp, gp, gX);

        if (io.gmuwts != nullptr && io.gflux[HBV_F_QSIM] != nullptr) {
            const float gm = __ldg(io.gflux[HBV_F_QSIM] + (int64_t)t * d.B + b) * Fl[HBV_F_QSIM];
            if (d.muwts_t_stride != 0) { if (valid) io.gmuwts[(int64_t)t * d.muwts_t_stride + lane] = gm; }
            else gmu_acc += gm;
        }
        if (io.gforcing != nullptr) {
            if constexpr (TR::HOURLY) { gX[0] *= d.inv_dt; gX[2] *= d.inv_dt; }
            float* gx = io.gforcing + ((int64_t)t * d.B + b) * d.nvar;
            if (shfl_reduce) {
                for (int o = DNM >> 1; o > 0; o >>= 1) {
                    gX[0] += __shfl_xor_sync(0xffffffffu, gX[0], o);
                    gX[1] += __shfl_xor_sync(0xffffffffu, gX[1], o);
                    gX[2] += __shfl_xor_sync(0xffffffffu, gX[2], o);
                }
                if (valid && j == 0) { gx[d.i_prcp] = gX[0]; gx[d.i_tmean] = gX[1]; gx[d.i_pet] = gX[2]; }
            }
        }

        // parameter gradients of this step.  Row T-1 also holds the static-parameter and routing
        // gradients (written by other code): it is written element-wise, not as a whole run.
        const bool direct = (t == d.T - 1);
        float* const orow = reinterpret_cast<float*>(lane_o + (t & 1) * g.obytes);
        float* const grow = gdyn_lane + (int64_t)t * (p_tstride / 4);
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (DS::is_dyn(i, 0)) {
                const float v = gp[i] * dpd[i];
                if (direct) { if (valid) grow[dense_col<NPAR, DM, LAYOUT>(i)] = v; }
                else if (valid) orow[dense_col<NPAR, DM, LAYOUT>(i)] = v;
            } else {
                gacc[i] += gp[i];
            }
        }

        if (tid == 0) bulk_wait_read<0>();    // the previous step's store has read the other buffer
        fence_proxy_async();
        __syncthreads();                      // staging buffer complete; slot `sl` consumed by all
        if (!direct) {
            const unsigned char* ob = obuf + (t & 1) * g.obytes;
            char* gdst = grun + t * p_tstride;
            if (tid == 0) {
                if (core) { bulk_s2g(gdst + head, ob + shg + head, core); }
                bulk_commit();
            }
            // <= 3 floats on either side of the aligned core
            if (tid < (int)(head >> 2) && (uint32_t)tid * 4 < pn)
                reinterpret_cast<float*>(gdst)[tid] = reinterpret_cast<const float*>(ob + shg)[tid];
            if (tid >= 4 && tid - 4 < (int)(tail >> 2)) {
                const uint32_t off = head + core + (uint32_t)(tid - 4) * 4;
                *reinterpret_cast<float*>(gdst + off) = *reinterpret_cast<const float*>(ob + shg + off);
            }
        }
        if (tid == 0) issue_step(t - NS, sl);
        shp = (shp - incp) & 15u;
        shf = (shf - incf) & 15u;
        shq = (shq - incq) & 15u;
        if (++sl == NS) { sl = 0; phase ^= 1u; }
    }
    if (tid == 0) bulk_wait<0>();             // stores complete before shared memory goes away
    // d(par)/d(raw) of the time-invariant parameters, recomputed here (as in hbv_bwd.cu)
    uint32_t lastmask = 0;
    resolve_params<NPAR, DM>(d, io.dyn, io.sta, nullptr, b, j, p, dpd, &lastmask);
    if (valid) {
        float* glast = gdyn_lane + (int64_t)(d.T - 1) * (p_tstride / 4);
#pragma unroll
        for (int i = 0; i < NPAR; ++i) {
            if (i < d.n_par && !DS::is_dyn(i, 0)) {
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

// ================================================================================================
// host: eligibility + launch
// ================================================================================================
// the dynamic columns must sit where the compiled layout expects them
template <int NPAR, int DM, int LAYOUT>
static bool layout_matches(const KDesc& d) {
    for (int i = 0; i < NPAR; ++i)
        if ((DM >> i) & 1)
            if (d.col[i] != dense_col<NPAR, DM, LAYOUT>(i)) return false;
    return true;
}

// HBV_B200_DENSE: 0 = always K1 / K2, 2 = take the dense kernels whenever the shapes allow it
// (tests), unset / 1 = where they win
static int dense_mode() {
    const long long v = opt(OPT_DENSE);
    return (v >= 0 && v <= 2) ? (int)v : 1;
}

// Common gate: compile-time dynamic set (dm > 0, no dropout), nmul 16, the dynamic columns are at
// least half of a row.  By default additionally: at least 8 time-varying parameters (>= 512 B of
// parameter bytes per basin-step, the HBM-bound regime) on a grid that fills every resident CTA
// slot of the GPU.  Measured on B200: with 3 dynamic parameters (hbv_2 family, 22.5k basins) K2
// beats K2d (4.7 vs 5.8 ms) — the per-step barrier and mbarrier round trip outweigh the saved
// load instructions — and on small grids (BASELINE configs[3], 2,500 units) the recurrence is
// latency-bound, where K2d's per-step synchronisation costs 45.7 vs 30.9 ms.
static bool dense_shape_ok(const KDesc& d, int dm) {
    const int mode = dense_mode();
    if (mode == 0 || dm <= 0 || d.nmul != DNM) return false;
    const int ndyn = popc_c((unsigned)dm);
    if (2 * ndyn * DNM < d.dyn_ncol) return false;
    const long long lanes = (long long)d.B * DNM;
    if (mode == 2) return lanes > 148LL * 4 * 32 * 2;
    return ndyn >= 8 && lanes >= 148LL * 16 * 32;
}

// 0 when every staged run of every CTA and time step starts on a 16 B boundary (then no slot
// needs room for the shift), else 16
static int run_slack(const KDesc& d, const void* a, const void* b, const void* c, const void* e) {
    const bool aligned = d.B % 4 == 0 && (d.BPB * d.dyn_ncol) % 4 == 0 && (d.BPB * d.nvar) % 4 == 0 &&
                         d.BPB % 4 == 0 &&
                         (reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                          reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(e)) % 16 == 0;
    return aligned ? 0 : 16;
}

static size_t smem_budget_4cta() { return (size_t)(227 * 1024) / 4 - 1024; }

template <typename K>
static int optin_smem(K k, size_t smem, std::atomic<int>* flag) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < HBV_MAX_DEVICES && flag[dev].load(std::memory_order_acquire) == 0) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        flag[dev].store(1, std::memory_order_release);
    }
    (void)smem;
    return 0;
}

static int opt_int(Opt o, int dflt) {
    const long long v = opt(o);
    return v >= 0 ? (int)v : dflt;
}

template <int VAR, bool BETAET, int DM, int LAYOUT>
static int launch_fwd_dense(KDesc d, const FwdPtrs& io, cudaStream_t st) {
    d.BPB = 128 / DNM;
    d.slack = run_slack(d, io.dyn, io.forcing, nullptr, nullptr);
    const DenseGeom g = dense_geom(d);
    const size_t fixed = 64 + g.tile_bytes;
    const size_t slot = g.pbytes + g.fbytes;
    int ns = opt_int(OPT_DENSE_NS, 0);
    if (ns <= 0) {
        ns = 8;                                   // deepest ring that still lets 4 CTAs share an SM
        while (ns > 4 && fixed + ns * slot > smem_budget_4cta()) ns -= DTC;
    }
    ns = (ns / DTC) * DTC;
    if (ns < 2 * DTC || ns > 8 * DTC || fixed + ns * slot > 200 * 1024) return HBV_NOT_ELIGIBLE;
    d.nstage = ns;
    auto k = hbv_fwd_dense_kernel<VAR, BETAET, DM, LAYOUT>;
    static std::atomic<int> optin[HBV_MAX_DEVICES];
    int rc = optin_smem(k, fixed + ns * slot, optin);
    if (rc) return rc;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    k<<<grid, d.BPB * DNM, fixed + ns * slot, st>>>(d, io);
    count_launch();
    count_dense_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET, int DM, int LAYOUT, int MINB>
static int launch_bwd_dense_m(KDesc d, const BwdPtrs& io, cudaStream_t st) {
    d.BPB = 128 / DNM;
    d.slack = run_slack(d, io.dyn, io.forcing, io.gdyn, io.gflux[HBV_F_QSIM]);
    const DenseGeom g = dense_geom(d);
    const size_t fixed = 64 + 2 * (size_t)g.obytes;
    const size_t slot = g.pbytes + g.fbytes + g.gbytes + g.sbytes;
    int ns = opt_int(OPT_DENSE_NS_BWD, 0);
    if (ns <= 0) {
        ns = 6;
        const size_t budget = (size_t)(227 * 1024) / MINB - 1024;
        while (ns > 2 && fixed + ns * slot > budget) --ns;
    }
    if (ns < 2 || ns > 8 || fixed + ns * slot > 200 * 1024) return HBV_NOT_ELIGIBLE;
    d.nstage = ns;
    auto k = hbv_bwd_dense_kernel<VAR, BETAET, DM, LAYOUT, MINB>;
    static std::atomic<int> optin[HBV_MAX_DEVICES];
    int rc = optin_smem(k, fixed + ns * slot, optin);
    if (rc) return rc;
    const int grid = (d.B + d.BPB - 1) / d.BPB;
    k<<<grid, d.BPB * DNM, fixed + ns * slot, st>>>(d, io);
    count_launch();
    count_dense_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

template <int VAR, bool BETAET, int DM, int LAYOUT>
static int launch_bwd_dense(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    const int m = opt_int(OPT_DENSE_MINB, 5);
    if (m == 6) return launch_bwd_dense_m<VAR, BETAET, DM, LAYOUT, 6>(d, io, st);
    if (m == 4) return launch_bwd_dense_m<VAR, BETAET, DM, LAYOUT, 4>(d, io, st);
    return launch_bwd_dense_m<VAR, BETAET, DM, LAYOUT, 5>(d, io, st);
}

// Try the dense forward.  Returns HBV_NOT_ELIGIBLE when the call is not eligible (the caller then takes K1).
template <int VAR, bool BETAET, int DM>
int try_fwd_dense(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if (!write_flux || io.drop != nullptr || !dense_shape_ok(d, DM)) return HBV_NOT_ELIGIBLE;
    if (d.ck_layout != 0 && io.ckpt != nullptr) return HBV_NOT_ELIGIBLE;      // (bulk copies of CTA-wide state rows: plane layout only)
    if (layout_matches<NPAR, DM, 0>(d)) return launch_fwd_dense<VAR, BETAET, DM, 0>(d, io, st);
    if (layout_matches<NPAR, DM, 1>(d)) return launch_fwd_dense<VAR, BETAET, DM, 1>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

// Try the dense adjoint: every-state-stored sweeps only, gradient rows with a time-invariant
// 16 B phase (B * ncol % 4 == 0) so a step's rows are one aligned bulk store.
template <int VAR, bool BETAET, int DM>
int try_bwd_dense(const KDesc& d, const BwdPtrs& io, cudaStream_t st) {
    constexpr int NPAR = Traits<VAR>::NPAR;
    if (d.K != 1 || d.ck_layout != 0 || io.drop != nullptr || io.gdyn == nullptr || !dense_shape_ok(d, DM)) return HBV_NOT_ELIGIBLE;
    if (reinterpret_cast<uintptr_t>(io.ckpt) % 16 != 0) return HBV_NOT_ELIGIBLE;
    if (((long long)d.B * d.dyn_ncol) % 4 != 0) return HBV_NOT_ELIGIBLE;
    if (layout_matches<NPAR, DM, 0>(d)) return launch_bwd_dense<VAR, BETAET, DM, 0>(d, io, st);
    if (layout_matches<NPAR, DM, 1>(d)) return launch_bwd_dense<VAR, BETAET, DM, 1>(d, io, st);
    return HBV_NOT_ELIGIBLE;
}

// the compiled (variant, dynamic set) pairs — the same sets hbv_fwd.cu / hbv_bwd.cu specialise
template int try_fwd_dense<HBV_VARIANT_HBV11P, true, DM_ALL14>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_dense<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_fwd_dense<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const FwdPtrs&, bool, cudaStream_t);
template int try_bwd_dense<HBV_VARIANT_HBV11P, true, DM_ALL14>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_dense<HBV_VARIANT_HBV2, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);
template int try_bwd_dense<HBV_VARIANT_HOURLY, true, DM_D3>(const KDesc&, const BwdPtrs&, cudaStream_t);

}  // namespace hbv
