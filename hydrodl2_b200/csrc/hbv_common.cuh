// hbv_common.cuh — kernel-side descriptor, parameter addressing and error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hbv_b200.h"
#include "hbv_step.cuh"

namespace hbv {

constexpr int HBV_MAX_DEVICES = 64;

// Device-side copy of hbv_desc_t with derived quantities (passed by value as a kernel argument).
struct KDesc {
    int T, B, nmul, n_par;
    int nvar, i_prcp, i_tmean, i_pet;
    int dyn_ncol, sta_ncol;
    int apply_sigmoid;
    int K;               // checkpoint interval
    int ck_layout;       // stored states: 0 = planes [segment][5][lane]; 1 = warp-major
                         // [lane / 32][segment][5][32] (hbv_desc_t.ckpt_layout)
    int muwts_t_stride;
    int BPB;             // basins per CTA
    int nstage;          // input ring: cp.async groups in flight + 1 (2..4); hbv_dense.cu: ring slots
    int slack;           // hbv_dense.cu: 16 when a staged run may start off a 16 B boundary, else 0;
                         // hbv_bwd.cu: offset (floats) of the forcing-gradient slab in shared memory
    float nearzero, dt, inv_dt;
    int src[HBV_MAX_PAR];
    int col[HBV_MAX_PAR];
    float lo[HBV_MAX_PAR];
    float span[HBV_MAX_PAR];  // hi - lo
};

// A translation unit with many explicit instantiations can be compiled as several objects
// (-DHBV_TU_PART=k, hydrodl2_b200/_build.py); without the macro it emits all of them.
#ifndef HBV_TU_PART
#define HBV_TU_PART 0
#endif
#define HBV_IN_PART(k) (HBV_TU_PART == 0 || HBV_TU_PART == (k))

// Stored-state addressing (floats): state s of segment g for global lane L sits at
//   ck_base(d, L) + (g * 5 + s) * ck_plane(d).
// Layout 1 gives a warp ONE contiguous 640 B run per stored step (and consecutive steps adjacent)
// instead of five 128 B rows 4 * nlane bytes apart: measured on BASELINE config 4's per-GPU grid,
// K1s 8.02 -> 7.82 ms and K2s 11.69 -> 11.30 ms.
__host__ __device__ __forceinline__ int64_t ck_nseg(const KDesc& d) { return d.K > 0 ? (d.T + d.K - 1) / d.K : 0; }
__host__ __device__ __forceinline__ int64_t ck_plane(const KDesc& d) {
    return d.ck_layout ? 32 : (int64_t)d.B * d.nmul;
}
__host__ __device__ __forceinline__ int64_t ck_base(const KDesc& d, int64_t lane) {
    return d.ck_layout ? (lane >> 5) * (ck_nseg(d) * 5 * 32) + (lane & 31) : lane;
}

struct FwdPtrs {
    const float* forcing; const float* dyn; const float* sta; const uint8_t* drop;
    const float* attrs; const float* muwts; const float* state_in;
    float* state_out; float* flux[HBV_MAX_FLUX]; float* state_series; float* ckpt;
};

struct BwdPtrs {
    const float* forcing; const float* dyn; const float* sta; const uint8_t* drop;
    const float* attrs; const float* muwts; const float* ckpt;
    const float* gflux[HBV_MAX_FLUX]; const float* gstate_out; const float* gstate_series;
    float* gdyn; float* gsta; float* gstate_in; float* gforcing; float* gmuwts;
    int zero_fill;
    int rows_before;     // rows in front of gdyn the adjoint zeroes as well (K2p only; else memset by the dispatcher)
};

// value in [0,1] (or raw) -> physical parameter
__device__ __forceinline__ float descale(const KDesc& d, int i, float raw) {
    const float v = d.apply_sigmoid ? (i == HBV_P_TT ? sigmoidf_(raw) : sigmoid_sfu(raw)) : raw;
    return v * d.span[i] + d.lo[i];
}

// d(par)/d(raw) given the physical value is not needed: recompute from raw.
__device__ __forceinline__ float descale_grad(const KDesc& d, int i, float raw) {
    if (d.apply_sigmoid) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        return d.span[i] * s * (1.0f - s);
    }
    return d.span[i];
}

// value + derivative with one sigmoid evaluation
__device__ __forceinline__ void descale_both(const KDesc& d, int i, float raw, float& val, float& dval) {
    if (d.apply_sigmoid) {
        const float s = (i == HBV_P_TT) ? sigmoidf_(raw) : sigmoid_sfu(raw);
        val = s * d.span[i] + d.lo[i];
        dval = d.span[i] * s * (1.0f - s);
    } else {
        val = raw * d.span[i] + d.lo[i];
        dval = d.span[i];
    }
}

// ---- which parameters are time-varying -----------------------------------------------------
// DM >= 0: compile-time bit mask (bit i = parameter i is read from row t of `dyn`); requires
//          no dropout mask.  DM = -1: runtime mask (any set, dropout allowed).
constexpr int DM_D2 = (1 << HBV_P_BETA) | (1 << HBV_P_BETAET);   // the reference's shipped set
constexpr int DM_D3 = DM_D2 | (1 << HBV_P_K0);                   // BASELINE.json configs[3] (hourly)
constexpr int DM_ALL14 = (1 << 14) - 1;                          // hbv_1_1p, every parameter dynamic (configs[2])

__host__ __device__ constexpr int popc_c(unsigned x) {
    int n = 0;
    while (x) { n += (int)(x & 1u); x >>= 1; }
    return n;
}

template <int NPAR, int DM>
struct DynSet {
    static constexpr bool STATIC = DM >= 0;
    static constexpr int NDYN = STATIC ? popc_c((unsigned)(DM < 0 ? 0 : DM)) : NPAR;
    static constexpr int NS = NDYN > 0 ? NDYN : 1;     // raw values carried per prefetched step
    // steps per prefetch buffer: two buffers (A/B) alternate, so inputs are requested
    // CL..2*CL steps before they are consumed
    static constexpr int CL = (NDYN <= 3) ? 2 : 1;
    __host__ __device__ static constexpr int slot(int i) {
        return STATIC ? popc_c((unsigned)(DM < 0 ? 0 : DM) & ((1u << i) - 1u)) : i;
    }
    __device__ __forceinline__ static bool is_dyn(int i, uint32_t dynmask) {
        return STATIC ? (((DM < 0 ? 0 : DM) >> i) & 1) : ((dynmask >> i) & 1u);
    }
};

// inputs of one time step as loaded (raw, not yet descaled)
template <int NS>
struct StepIn { float P, T, PET; float raw[NS]; };

// Resolve every parameter's source for this lane; load + descale the time-invariant ones.
// Returns the lane's dynamic mask.  dpd / lastmask may be null (forward).
template <int NPAR, int DM>
__device__ __forceinline__ uint32_t resolve_params(const KDesc& d, const float* dyn, const float* sta,
                                                   const uint8_t* drop, int b, int j, float (&p)[NPAR],
                                                   float* dpd, uint32_t* lastmask) {
    uint32_t dynmask = 0;
    const float* dyn_last = dyn + ((int64_t)(d.T - 1) * d.B + b) * d.dyn_ncol + j;
#pragma unroll
    for (int i = 0; i < NPAR; ++i) {
        p[i] = 0.f;
        if (dpd) dpd[i] = 0.f;
        if (i < d.n_par) {
            int src = d.src[i];
            if (DM < 0 && src == HBV_SRC_DYN_T && drop != nullptr && drop[(int64_t)i * d.B + b]) src = HBV_SRC_DYN_LAST;
            const bool isdyn = (DM >= 0) ? (((DM < 0 ? 0 : DM) >> i) & 1) : (src == HBV_SRC_DYN_T);
            if (isdyn) {
                dynmask |= (1u << i);
            } else {
                float raw;
                if (src == HBV_SRC_STA) raw = __ldg(sta + (int64_t)b * d.sta_ncol + d.col[i] + j);
                else { raw = __ldg(dyn_last + d.col[i]); if (lastmask) *lastmask |= (1u << i); }
                float v, dv;
                descale_both(d, i, raw, v, dv);
                p[i] = v;
                if (dpd) dpd[i] = dv;
            }
        }
    }
    return dynmask;
}

template <int NPAR, int DM>
__device__ __forceinline__ void load_step(const KDesc& d, const float* fptr, int64_t f_tstride,
                                          const float* dyn_lane, int64_t dyn_tstride, uint32_t dynmask,
                                          int t, StepIn<DynSet<NPAR, DM>::NS>& in) {
    using DS = DynSet<NPAR, DM>;
    const float* fr = fptr + (int64_t)t * f_tstride;
    in.P = __ldg(fr + d.i_prcp);
    in.T = __ldg(fr + d.i_tmean);
    in.PET = __ldg(fr + d.i_pet);
    if (DS::STATIC ? (DM != 0) : (dynmask != 0)) {
        const float* dr = dyn_lane + (int64_t)t * dyn_tstride;
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, dynmask)) in.raw[DS::slot(i)] = __ldg(dr + d.col[i]);
    }
}

template <int NPAR, int DM>
__device__ __forceinline__ void apply_dyn(const KDesc& d, uint32_t dynmask,
                                          const StepIn<DynSet<NPAR, DM>::NS>& in, float (&p)[NPAR], float* dpd) {
    using DS = DynSet<NPAR, DM>;
#pragma unroll
    for (int i = 0; i < NPAR; ++i)
        if (DS::is_dyn(i, dynmask)) {
            if (dpd) { float v, dv; descale_both(d, i, in.raw[DS::slot(i)], v, dv); p[i] = v; dpd[i] = dv; }
            else p[i] = descale(d, i, in.raw[DS::slot(i)]);
        }
}

// ---- shared-memory input ring (cp.async) ----------------------------------------------------
// Each thread stages its OWN inputs (3 forcings + the raw values of its dynamic parameters) for
// the coming time steps with 4-byte cp.async copies (LDGSTS: no registers, no scoreboard slot),
// `depth = RC * nstage` steps deep, and later reads them back with immediate-offset LDS.  Layout
// [ring step][thread][NSP] with NSP odd: the per-thread stride is conflict-free and no thread
// ever reads another thread's slots, so the ring needs no barrier — only cp.async.wait_group.
// Slot 0..2 = P, T, PET; slot 3 + s = dynamic parameter (s = compact index for a compile-time
// set, the parameter index for the runtime set); the adjoint adds upstream-gradient slots.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(float* smem_dst, const float* gsrc) {     // 8 B aligned
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {    // 16 B aligned
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// wait until the oldest of `nstage - 1` outstanding groups has landed
__device__ __forceinline__ void ring_wait(int nstage) {
    if (nstage >= 4) cp_async_wait<2>();
    else if (nstage == 3) cp_async_wait<1>();
    else cp_async_wait<0>();
}

template <int NPAR, int DM>
struct RingSlots {
    using DS = DynSet<NPAR, DM>;
    static constexpr int NDS = DS::STATIC ? DS::NDYN : NPAR;     // parameter slots
    __host__ __device__ static constexpr int par(int i) { return 3 + (DS::STATIC ? DS::slot(i) : i); }
    static constexpr int FIRST_FREE = 3 + NDS;
};

// stage the inputs of one time step into `dst` (this thread's NSP floats).  frow / drow point at
// this thread's forcing row and dynamic-parameter row of that step; column offsets are 32-bit so
// each address is one IMAD.WIDE
template <int NPAR, int DM>
__device__ __forceinline__ void ring_issue_step(const KDesc& d, const float* frow, const float* drow,
                                                uint32_t dynmask, float* dst) {
    using DS = DynSet<NPAR, DM>;
    using RS = RingSlots<NPAR, DM>;
    cp_async4(dst + 0, frow + (unsigned)d.i_prcp);
    cp_async4(dst + 1, frow + (unsigned)d.i_tmean);
    cp_async4(dst + 2, frow + (unsigned)d.i_pet);
    if (DS::STATIC ? (DM != 0) : (dynmask != 0)) {
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, dynmask)) cp_async4(dst + RS::par(i), drow + (unsigned)d.col[i]);
    }
}

// The same inputs held in registers (throughput regime: many resident warps hide the latency,
// and plain LDG costs the LSU/MIO pipe a third of what LDGSTS + LDS do).  Same slot numbering.
template <int NPAR, int DM>
struct RegIn {
    float v[RingSlots<NPAR, DM>::FIRST_FREE];
    __device__ __forceinline__ float operator[](int k) const { return v[k]; }
};

template <int NPAR, int DM>
__device__ __forceinline__ void reg_load_step(const KDesc& d, const float* frow, const float* drow,
                                              uint32_t dynmask, RegIn<NPAR, DM>& in) {
    using DS = DynSet<NPAR, DM>;
    using RS = RingSlots<NPAR, DM>;
    in.v[0] = __ldg(frow + (unsigned)d.i_prcp);
    in.v[1] = __ldg(frow + (unsigned)d.i_tmean);
    in.v[2] = __ldg(frow + (unsigned)d.i_pet);
    if (DS::STATIC ? (DM != 0) : (dynmask != 0)) {
#pragma unroll
        for (int i = 0; i < NPAR; ++i)
            if (DS::is_dyn(i, dynmask)) in.v[RS::par(i)] = __ldg(drow + (unsigned)d.col[i]);
    }
}

// descale this step's dynamic parameters from the staged inputs (ring slot or registers)
template <int NPAR, int DM, typename IN>
__device__ __forceinline__ void ring_apply_dyn(const KDesc& d, uint32_t dynmask, const IN& in,
                                               float (&p)[NPAR], float* dpd) {
    using DS = DynSet<NPAR, DM>;
    using RS = RingSlots<NPAR, DM>;
#pragma unroll
    for (int i = 0; i < NPAR; ++i)
        if (DS::is_dyn(i, dynmask)) {
            if (dpd) { float v, dv; descale_both(d, i, in[RS::par(i)], v, dv); p[i] = v; dpd[i] = dv; }
            else p[i] = descale(d, i, in[RS::par(i)]);
        }
}

// Steps per cp.async group: a whole output chunk (4 steps) when a step is a few floats per thread,
// one step when it is many (runtime set, all-dynamic) so that a deep ring stays small.
template <int NPAR, int DM>
struct RingCfg {
    static constexpr int NSP = RingSlots<NPAR, DM>::FIRST_FREE | 1;   // floats / thread / step (odd)
    static constexpr int RC = (NSP <= 7) ? 4 : 1;
};

// host: number of groups (2..4) whose ring fits `budget` bytes next to `fixed` bytes of other
// shared memory: deep for small grids, where a CTA owns an SM sub-partition and only distance
// hides latency; shallower when many CTAs share an SM.  0 = does not fit.
inline int choose_nstage(int NT, int nsp, int rc, size_t fixed, long long grid, size_t* bytes) {
    const size_t budget = (grid <= 2LL * 148) ? 100 * 1024 : 56 * 1024;
    const size_t group = (size_t)rc * NT * nsp * sizeof(float);
    int n = (rc == 1) ? 4 : 3;
    while (n > 2 && fixed + n * group > budget) --n;
    *bytes = n * group;
    return (fixed + n * group <= 100 * 1024) ? n : 0;
}

// host: the runtime dynamic set as a bit mask, or -1 when a dropout mask forces the generic path
inline int static_dynmask(const KDesc& d, bool has_drop) {
    if (has_drop) return -1;
    int m = 0;
    for (int i = 0; i < d.n_par; ++i) if (d.src[i] == HBV_SRC_DYN_T) m |= (1 << i);
    return m;
}

// Compiler-level ordering point.  The prefetch loads of a later step are placed AFTER the point
// where the current step's inputs are consumed: outstanding loads share the warp's six
// scoreboard slots, and a wait on a slot waits for every load armed on it — if the new loads were
// issued first, consuming the old values would stall for a full HBM round trip every step.
__device__ __forceinline__ void order_point() { asm volatile("" ::: "memory"); }

// ---- output staging tile of the forward kernels (nmul reduction through shared memory) --------
constexpr int NFP = 12;      // floats per lane per step in the staging tile (3 x float4)
__host__ __device__ inline int tile_bstride(int nmul) {
    // per-basin stride in floats; +12 keeps 128-bit accesses conflict-free for nmul = 16
    return nmul * NFP + 12;
}

// ---- TMA-staged kernels for dense-dynamic runs (hbv_dense.cu) ---------------------------------
// Return HBV_NOT_ELIGIBLE when the call does not fit them; the caller then launches K1 / K2.
constexpr int HBV_NOT_ELIGIBLE = -1000;
template <int VAR, bool BETAET, int DM>
int try_fwd_dense(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st);
template <int VAR, bool BETAET, int DM>
int try_bwd_dense(const KDesc& d, const BwdPtrs& io, cudaStream_t st);

// ---- lean kernels for the standard layout in the throughput regime (hbv_lean.cu) ---------------
template <int VAR, bool BETAET, int DM>
int try_fwd_lean(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st);
template <int VAR, bool BETAET, int DM>
int try_bwd_lean(const KDesc& d, const BwdPtrs& io, cudaStream_t st);
template <int VAR, bool BETAET>
int try_fwd_lean_warm(const KDesc& d, const FwdPtrs& io, cudaStream_t st);

// ---- stage-pipelined kernels for the latency-bound regime (hbv_pipe.cu) ------------------------
template <int VAR, bool BETAET, int DM>
int try_fwd_pipe(const KDesc& d, const FwdPtrs& io, bool write_flux, cudaStream_t st);
template <int VAR, bool BETAET, int DM>
int try_bwd_pipe(const KDesc& d, const BwdPtrs& io, cudaStream_t st);
template <int VAR, bool BETAET>
int try_fwd_pipe_warm(const KDesc& d, const FwdPtrs& io, cudaStream_t st);

// ---- experiment switches -----------------------------------------------------------------------
// Read from the environment ONCE, when the library is first used (HBV_B200_<NAME>), and settable
// at run time through hbv_b200_set_option("<name>", value) — no getenv on the dispatch path.
// -1 = unset (the library's own policy decides).
enum Opt {
    OPT_LEAN = 0,         // 0: never the standard-layout kernels K1s / K2s (nor K1p / K2p)
    OPT_PIPE,             // 0: never the stage-pipelined kernels K1p / K2p
    OPT_PIPE_MAX,         // largest grid (lanes) K1p / K2p are used for
    OPT_RING,             // 0 / 1: force the register / cp.async-ring input path of K1 / K2
    OPT_LEAN_SMALL,       // grid size (lanes) up to which K1s uses its ring form
    OPT_LEAN_BWD_RING,    // 0: K2s register form
    OPT_LEAN_DEEP,        // grid size (lanes) up to which K1s' 128-thread form takes its inputs through the chunk ring
    OPT_DENSE,            // 0: never K1d / K2d, 2: wherever the shapes allow (tests), 1 / unset: where they win
    OPT_DENSE_NS, OPT_DENSE_NS_BWD, OPT_DENSE_MINB,
    OPT_CKPT,             // checkpoint interval hbv_b200_auto_ckpt returns (experiments)
    OPT_ADJ_BPB,          // basins per CTA of K3's forward (experiments; unset: by measurement)
    OPT_COPY_BLOCKS,      // hbv_b200_copy_cols: 256-thread blocks per SM (unset: 2)
    OPT_CKPT_LAYOUT,      // read by the host side (ops.py): 0 / 1 force the stored-state layout (unset: policy)
    OPT_COUNT
};
long long opt(Opt o);

void set_error(const char* msg);
void count_launch(int n = 1);
void count_dense_launch();
void count_lean_launch();
void count_pipe_launch();

}  // namespace hbv
