// pair_route.cu — K4': distributed (gage, unit) pair routing of hbv_2_hourly, forward + adjoint.
//
// Replaces Hbv_2_hourly.distr_routing / _frac_shift1d (models/hbv/hbv_2_hourly.py:800-897): the
// reference gathers the area-weighted runoff of every (gage, unit) pair into a [T, n_pairs]
// tensor, builds a 72-tap gamma UH per pair (uh_gamma), shifts it by a fractional lag
// (two torch.gather + masks), convolves with a grouped conv1d (groups = n_pairs), scatter_adds
// the pairs to their gages and divides by the upstream area.  Here:
//   pair_uh        one thread per pair: normalised gamma pdf + fractional shift -> uh[M][P]
//   pair_conv      shared-memory-staged convolution, one CTA = 32 pairs x 64 time steps; the
//                  gather x[t][p] = Qs[t][unit(p)] * area(unit(p)) is fused into the tile load
//   seg_sum        deterministic segmented sum (CSR by gage / CSC by unit) + scaling
//   pair_conv_bwd  adjoint conv (gradient gather g[t][gage(p)] / denom fused into the tile load)
//                  + per-tap dot products for d/dUH, reduced over time chunks
//   pair_uh_bwd    d/dUH -> (route_a, route_b, route_tau) through the shift and the gamma pdf
// The same conv kernels serve the plain per-unit routing of the hourly model (lenF = 72, identity
// gather), which the 16-tap register-window kernel of uh_route.cu cannot hold.
#include "hbv_common.cuh"

namespace hbv {

constexpr int PM = 128;       // max taps
constexpr int PTT = 64;       // time steps per tile
constexpr int PPB = 32;       // pairs per CTA (one per lane)
constexpr int PTY = 8;        // time threads per pair

struct PDesc {
    int T, P, M, lag_uh, tchunk, nchunk;
    float a_lo, a_span, b_lo, b_span, t_lo, t_span;
};

__device__ __forceinline__ float gamma_w(float t, float aa, float th) {
    return powf(t, aa - 1.f) * expf(-t / th);
}

// uh[k][p] = fractional shift of the normalised gamma pdf (uh_routing.py:5-22 +
// hbv_2_hourly.py:858-897).  par: [P, 3] in [0, 1] (route_a, route_b, route_tau).
__global__ void pair_uh_kernel(const PDesc d, const float* __restrict__ par, float* __restrict__ uh) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    const float a = par[p * 3 + 0] * d.a_span + d.a_lo;
    const float b = par[p * 3 + 1] * d.b_span + d.b_lo;
    const float aa = fmaxf(a, 0.f) + 0.1f;
    const float th = fmaxf(b, 0.f) + 0.5f;
    float sum = 0.f;
    for (int k = 0; k < d.M; ++k) sum += gamma_w((float)k + 0.5f, aa, th);
    const float inv = 1.f / sum;
    int kk = 0;
    float f = 0.f;
    if (d.lag_uh) {
        const float tau = par[p * 3 + 2] * d.t_span + d.t_lo;
        const float fl = floorf(tau);
        kk = (int)fl;
        f = tau - fl;
    }
    for (int k = 0; k < d.M; ++k) {
        const int i0 = k - kk, i1 = k - kk - 1;
        const float w0 = (i0 >= 0 && i0 < d.M) ? gamma_w((float)i0 + 0.5f, aa, th) * inv : 0.f;
        const float w1 = (i1 >= 0 && i1 < d.M) ? gamma_w((float)i1 + 0.5f, aa, th) * inv : 0.f;
        uh[(int64_t)k * d.P + p] = d.lag_uh ? ((1.f - f) * w0 + f * w1) : w0;
    }
}

// y[t][p] = sum_k uh[k][p] * x[t-k][p],  x[t][p] = src[t][col[p]] * scale[col[p]]
// REV: y[t][p] = sum_k uh[k][p] * x[t+k][p] (adjoint)
template <bool REV>
__global__ void __launch_bounds__(PPB * PTY)
pair_conv_kernel(const PDesc d, const float* __restrict__ uh, const float* __restrict__ src, int src_stride,
                 const int* __restrict__ col, const float* __restrict__ scale, float* __restrict__ y) {
    extern __shared__ float sm[];
    float* us = sm;                       // [M][PPB]
    float* xs = sm + d.M * PPB;           // [PTT + M - 1][PPB]
    const int px = threadIdx.x % PPB;
    const int ty = threadIdx.x / PPB;
    const int p = blockIdx.x * PPB + px;
    const int t0 = blockIdx.y * PTT;
    const bool pv = p < d.P;
    int c = 0;
    float sc = 0.f;
    if (pv) { c = col ? col[p] : p; sc = scale ? scale[c] : 1.f; }
    for (int k = ty; k < d.M; k += PTY) us[k * PPB + px] = pv ? uh[(int64_t)k * d.P + p] : 0.f;
    const int nrow = PTT + d.M - 1;
    // window rows r = 0..nrow-1 map to time  t0 - (M-1) + r  (forward)  /  t0 + r  (reverse)
    for (int r = ty; r < nrow; r += PTY) {
        const int t = REV ? (t0 + r) : (t0 - (d.M - 1) + r);
        xs[r * PPB + px] = (pv && t >= 0 && t < d.T) ? src[(int64_t)t * src_stride + c] * sc : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < PTT; i += PTY) {
        const int t = t0 + i;
        if (t >= d.T) break;
        float acc = 0.f;
        if (REV) {
            for (int k = 0; k < d.M; ++k) acc = fmaf(us[k * PPB + px], xs[(i + k) * PPB + px], acc);
        } else {
            for (int k = 0; k < d.M; ++k) acc = fmaf(us[k * PPB + px], xs[(i + d.M - 1 - k) * PPB + px], acc);
        }
        if (pv) y[(int64_t)t * d.P + p] = acc;
    }
}

// out[t][s] = mul[s] * sum_{i in [off[s], off[s+1])} val[t][perm ? perm[i] : i]
__global__ void seg_sum_kernel(int T, int S, int P, const int* __restrict__ off, const int* __restrict__ perm,
                               const float* __restrict__ val, const float* __restrict__ mul,
                               float* __restrict__ out, int out_stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (t >= T) return;
    float acc = 0.f;
    const float* row = val + (int64_t)t * P;
    for (int i = off[s]; i < off[s + 1]; ++i) acc += row[perm ? perm[i] : i];
    out[(int64_t)t * out_stride + s] = acc * (mul ? mul[s] : 1.f);
}

// dU[k][chunk][p] = sum_{t in chunk} g[t][p] * x[t-k][p]
//   g[t][p] = gsrc[t][gcol[p]] * gscale[gcol[p]],  x[t][p] = src[t][col[p]] * scale[col[p]]
__global__ void __launch_bounds__(PPB * PTY)
pair_duh_kernel(const PDesc d, const float* __restrict__ gsrc, int g_stride, const int* __restrict__ gcol,
                const float* __restrict__ gscale, const float* __restrict__ src, int src_stride,
                const int* __restrict__ col, const float* __restrict__ scale, float* __restrict__ ws) {
    extern __shared__ float sm[];
    float* gs = sm;                        // [PTT][PPB]
    float* xs = sm + PTT * PPB;            // [PTT + M - 1][PPB]
    const int px = threadIdx.x % PPB;
    const int ty = threadIdx.x / PPB;
    const int p = blockIdx.x * PPB + px;
    const int ch = blockIdx.y;
    const bool pv = p < d.P;
    int c = 0, gc = 0;
    float sc = 0.f, gsc = 0.f;
    if (pv) {
        c = col ? col[p] : p; sc = scale ? scale[c] : 1.f;
        gc = gcol ? gcol[p] : p; gsc = gscale ? gscale[gc] : 1.f;
    }
    constexpr int KPT = PM / PTY;          // taps per thread: k = ty + PTY * q
    float acc[KPT];
#pragma unroll
    for (int q = 0; q < KPT; ++q) acc[q] = 0.f;
    const int tbeg = ch * d.tchunk;
    const int tend = min(d.T, tbeg + d.tchunk);
    const int nrow = PTT + d.M - 1;
    for (int t0 = tbeg; t0 < tend; t0 += PTT) {
        __syncthreads();
        for (int i = ty; i < PTT; i += PTY) {
            const int t = t0 + i;
            gs[i * PPB + px] = (pv && t < tend) ? gsrc[(int64_t)t * g_stride + gc] * gsc : 0.f;
        }
        for (int r = ty; r < nrow; r += PTY) {
            const int t = t0 - (d.M - 1) + r;
            xs[r * PPB + px] = (pv && t >= 0 && t < d.T) ? src[(int64_t)t * src_stride + c] * sc : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
            const int k = ty + PTY * q;
            if (k < d.M) {
                float a = acc[q];
                for (int i = 0; i < PTT; ++i) a = fmaf(gs[i * PPB + px], xs[(i + d.M - 1 - k) * PPB + px], a);
                acc[q] = a;
            }
        }
    }
    if (pv) {
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
            const int k = ty + PTY * q;
            if (k < d.M) ws[((int64_t)k * d.nchunk + ch) * d.P + p] = acc[q];
        }
    }
}

// d/dUH -> d/d(par) through the fractional shift and the normalised gamma pdf
__global__ void pair_uh_bwd_kernel(const PDesc d, const float* __restrict__ par, const float* __restrict__ ws,
                                   float* __restrict__ gpar) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    const float a = par[p * 3 + 0] * d.a_span + d.a_lo;
    const float b = par[p * 3 + 1] * d.b_span + d.b_lo;
    const float aa = fmaxf(a, 0.f) + 0.1f;
    const float th = fmaxf(b, 0.f) + 0.5f;
    float sum = 0.f;
    for (int k = 0; k < d.M; ++k) sum += gamma_w((float)k + 0.5f, aa, th);
    const float inv = 1.f / sum;
    int kk = 0;
    float f = 0.f;
    if (d.lag_uh) {
        const float tau = par[p * 3 + 2] * d.t_span + d.t_lo;
        const float fl = floorf(tau);
        kk = (int)fl;
        f = tau - fl;
    }
    // g_y[k] = dL/d(shifted uh[k]);  g_u[j] = dL/d(normalised w[j]) = (1-f) g_y[j+kk] + f g_y[j+kk+1]
    float gf = 0.f, mL = 0.f, mT = 0.f, sL = 0.f, sT = 0.f, s0 = 0.f;
    for (int j = 0; j < d.M; ++j) {
        const float t = (float)j + 0.5f;
        const float u = gamma_w(t, aa, th) * inv;
        float gy0 = 0.f, gy1 = 0.f;
        const int k0 = j + kk, k1 = j + kk + 1;
        if (k0 < d.M) for (int c = 0; c < d.nchunk; ++c) gy0 += ws[((int64_t)k0 * d.nchunk + c) * d.P + p];
        if (d.lag_uh && k1 < d.M) for (int c = 0; c < d.nchunk; ++c) gy1 += ws[((int64_t)k1 * d.nchunk + c) * d.P + p];
        const float gu = d.lag_uh ? ((1.f - f) * gy0 + f * gy1) : gy0;
        gf += u * (gy1 - gy0);                 // d/df of (1-f) u[k-kk] + f u[k-kk-1]
        const float L = logf(t);
        mL += u * L; mT += u * t;
        sL += gu * u * L; sT += gu * u * t; s0 += gu * u;
    }
    const float gaa = sL - s0 * mL;
    const float gth = (sT - s0 * mT) / (th * th);
    gpar[p * 3 + 0] = (a > 0.f ? gaa : 0.f) * d.a_span;
    gpar[p * 3 + 1] = (b > 0.f ? gth : 0.f) * d.b_span;
    gpar[p * 3 + 2] = d.lag_uh ? gf * d.t_span : 0.f;
}

static int make_pdesc(const hbv_pair_desc_t* s, PDesc& d) {
    if (!s) { set_error("null pair desc"); return HBV_E_NULL; }
    if (s->abi_version != HBV_B200_ABI_VERSION) { set_error("ABI version mismatch"); return HBV_E_ABI; }
    if (s->T <= 0 || s->n_pairs <= 0 || s->lenF <= 0 || s->n_units <= 0 || s->n_gages <= 0) { set_error("bad pair-routing shape"); return HBV_E_SHAPE; }
    d.T = s->T; d.P = s->n_pairs; d.M = s->lenF < s->T ? s->lenF : s->T;
    if (d.M > PM) { set_error("lenF > 128 not supported"); return HBV_E_SHAPE; }
    d.lag_uh = s->lag_uh;
    d.a_lo = s->a_lo; d.a_span = s->a_hi - s->a_lo; d.b_lo = s->b_lo; d.b_span = s->b_hi - s->b_lo;
    d.t_lo = s->tau_lo; d.t_span = s->tau_hi - s->tau_lo;
    d.nchunk = hbv_b200_pair_chunks(s->T);
    d.tchunk = ((s->T + d.nchunk - 1) / d.nchunk + PTT - 1) / PTT * PTT;
    return 0;
}

}  // namespace hbv

using namespace hbv;

extern "C" int hbv_b200_pair_chunks(int32_t T) {
    int n = (T + 2047) / 2048;
    return n < 1 ? 1 : n;
}

extern "C" int hbv_b200_pair_route_fwd(const hbv_pair_desc_t* desc, const float* par, const float* qs,
                                       const float* areas, const int32_t* pair_col,
                                       const int32_t* gage_off, const float* inv_denom, float* uh,
                                       float* lag, float* out, void* stream) {
    PDesc d;
    int rc = make_pdesc(desc, d);
    if (rc) return rc;
    if (!par || !qs || !gage_off || !uh || !lag || !out) { set_error("null pointer"); return HBV_E_NULL; }
    cudaStream_t st = (cudaStream_t)stream;
    pair_uh_kernel<<<(d.P + 127) / 128, 128, 0, st>>>(d, par, uh);
    dim3 grid((d.P + PPB - 1) / PPB, (d.T + PTT - 1) / PTT);
    const size_t smem = (size_t)(d.M + PTT + d.M - 1) * PPB * sizeof(float);
    pair_conv_kernel<false><<<grid, PPB * PTY, smem, st>>>(d, uh, qs, desc->n_units, pair_col, areas, lag);
    dim3 g2((d.T + 127) / 128, desc->n_gages);
    seg_sum_kernel<<<g2, 128, 0, st>>>(d.T, desc->n_gages, d.P, gage_off, nullptr, lag, inv_denom, out, desc->n_gages);
    count_launch(3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

extern "C" int hbv_b200_pair_route_bwd(const hbv_pair_desc_t* desc, const float* par, const float* qs,
                                       const float* areas, const int32_t* pair_col,
                                       const int32_t* pair_row, const float* inv_denom,
                                       const int32_t* unit_off, const int32_t* unit_perm,
                                       const float* uh, const float* g_out, float* g_lag_ws,
                                       float* duh_ws, float* g_qs, float* g_par, void* stream) {
    PDesc d;
    int rc = make_pdesc(desc, d);
    if (rc) return rc;
    if (!par || !qs || !pair_row || !unit_off || !unit_perm || !uh || !g_out || !g_lag_ws || !duh_ws || !g_qs || !g_par) {
        set_error("null pointer"); return HBV_E_NULL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((d.P + PPB - 1) / PPB, (d.T + PTT - 1) / PTT);
    const size_t smem = (size_t)(d.M + PTT + d.M - 1) * PPB * sizeof(float);
    // g_x[t][p] = sum_k uh[k][p] * g[t+k][p],  g[t][p] = g_out[t][row[p]] * inv_denom[row[p]]
    pair_conv_kernel<true><<<grid, PPB * PTY, smem, st>>>(d, uh, g_out, desc->n_gages, pair_row, inv_denom, g_lag_ws);
    // g_qs[t][u] = area[u] * sum_{pairs of unit u} g_x[t][p]
    dim3 g2((d.T + 127) / 128, desc->n_units);
    seg_sum_kernel<<<g2, 128, 0, st>>>(d.T, desc->n_units, d.P, unit_off, unit_perm, g_lag_ws, areas, g_qs, desc->n_units);
    dim3 g3((d.P + PPB - 1) / PPB, d.nchunk);
    const size_t smem3 = (size_t)(PTT + PTT + d.M - 1) * PPB * sizeof(float);
    pair_duh_kernel<<<g3, PPB * PTY, smem3, st>>>(d, g_out, desc->n_gages, pair_row, inv_denom, qs, desc->n_units,
                                                   pair_col, areas, duh_ws);
    pair_uh_bwd_kernel<<<(d.P + 127) / 128, 128, 0, st>>>(d, par, duh_ws, g_par);
    count_launch(4);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}
