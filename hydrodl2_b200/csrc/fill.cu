// fill.cu — thin zero-fill of the dense parameter-gradient plane.
//
// The autograd contract returns d/d(parameters) with the full [T, B, ncol] shape, most of it
// zeros.  On latency-bound grids the recurrence kernels leave HBM idle but own
// the schedulers: a full-occupancy memset next to them starves them for longer than the fill
// takes alone (round 1: a 76 us memset stretched C2's 88 us warm-up kernel to 275 us).  This
// kernel writes the zeros with almost no instruction issue: one warp per CTA, one CTA per SM (or
// fewer), each lane-0 thread streams `cp.async.bulk.global.shared::cta` stores (SASS UBLKCP.G.S)
// of a 16 KB zeroed shared-memory buffer — a few hundred instructions per CTA for hundreds of MB —
// so it can run on a side stream underneath the warm-up / forward kernels.
#include "hbv_common.cuh"

namespace hbv {

constexpr int FILL_CHUNK = 16 * 1024;

__global__ void __launch_bounds__(32, 1)
fill_zero_kernel(char* __restrict__ dst, int64_t nbytes) {
    __shared__ __align__(128) char z[FILL_CHUNK];
    for (int i = threadIdx.x; i < FILL_CHUNK / 16; i += 32) reinterpret_cast<uint4*>(z)[i] = make_uint4(0, 0, 0, 0);
    // the bulk store reads shared memory through the async proxy: order the generic-proxy writes first
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    // head (to the first 16 B boundary) and tail (after the last one) with plain byte stores
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    int64_t head = (int64_t)((16 - (a & 15)) & 15);
    if (head > nbytes) head = nbytes;
    const int64_t body = ((nbytes - head) / 16) * 16;
    if (blockIdx.x == 0) {
        for (int64_t i = threadIdx.x; i < head; i += 32) dst[i] = 0;
        for (int64_t i = head + body + threadIdx.x; i < nbytes; i += 32) dst[i] = 0;
    }
    if (threadIdx.x == 0) {
        char* base = dst + head;
        const int64_t nchunk = (body + FILL_CHUNK - 1) / FILL_CHUNK;
        const unsigned zs = (unsigned)__cvta_generic_to_shared(z);
        int inflight = 0;
        for (int64_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
            const int64_t off = c * FILL_CHUNK;
            const unsigned len = (unsigned)((body - off) < FILL_CHUNK ? (body - off) : FILL_CHUNK);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(base + off), "r"(zs), "r"(len) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) {       // bound the number of outstanding groups
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
                inflight = 4;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

}  // namespace hbv

extern "C" int hbv_b200_fill_zero(void* ptr, int64_t nbytes, int32_t n_ctas, void* stream) {
    using namespace hbv;
    if (nbytes < 0 || (nbytes > 0 && !ptr)) { set_error("fill_zero: bad arguments"); return HBV_E_NULL; }
    if (nbytes == 0) return 0;
    int grid = n_ctas > 0 ? n_ctas : 148;
    const int64_t nchunk = (nbytes + FILL_CHUNK - 1) / FILL_CHUNK;
    if (grid > nchunk) grid = (int)nchunk;
    fill_zero_kernel<<<grid, 32, 0, (cudaStream_t)stream>>>(static_cast<char*>(ptr), nbytes);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

// ---- column-block staging between pinned host memory and the device --------------------------
// The kernels read 2 of the 13 parameter blocks of a packed `parameters` tensor plus its last
// row (hbv.py:201-214), and 85 % of the dense gradient they return is zeros: copying whole tensors
// moves ~7x the bytes PCIe needs to carry.  copy_cols moves one [rows x ncols] float block that
// sits at column `col0` of a row-major [rows, row_stride] matrix, same layout on both sides,
// between a (mapped, pinned) host pointer and a device pointer — the GPU reads / writes the host
// memory directly over PCIe (UVA), `width`-byte runs, many requests in flight.  A DMA-engine
// 2-D copy of 64-byte-wide rows is descriptor-bound; this kernel is link-bound.
namespace hbv {

template <typename V>
__global__ void __launch_bounds__(256)
copy_cols_kernel(V* __restrict__ dst, const V* __restrict__ src, int64_t rows, int64_t row_stride_v,
                 int ncols_v) {
    // one thread per vector element of the block; consecutive threads walk a row's run
    const int64_t n = rows * ncols_v;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncols_v;
        const int c = (int)(i - r * ncols_v);
        dst[r * row_stride_v + c] = src[r * row_stride_v + c];
    }
}

}  // namespace hbv

extern "C" int hbv_b200_copy_cols(float* dst, const float* src, int64_t rows, int64_t row_stride,
                                  int32_t col0, int32_t ncols, void* stream) {
    using namespace hbv;
    if (!dst || !src || rows < 0 || row_stride <= 0 || col0 < 0 || ncols < 0 || col0 + ncols > row_stride) {
        set_error("copy_cols: bad arguments");
        return HBV_E_SHAPE;
    }
    if (rows == 0 || ncols == 0) return 0;
    float* d = dst + col0;
    const float* s = src + col0;
    const bool v4 = (reinterpret_cast<uintptr_t>(d) % 16 == 0) && (reinterpret_cast<uintptr_t>(s) % 16 == 0) &&
                    row_stride % 4 == 0 && ncols % 4 == 0;
    const bool v2 = (reinterpret_cast<uintptr_t>(d) % 8 == 0) && (reinterpret_cast<uintptr_t>(s) % 8 == 0) &&
                    row_stride % 2 == 0 && ncols % 2 == 0;
    const int vec = v4 ? 4 : (v2 ? 2 : 1);
    const int64_t n = rows * (ncols / vec);
    // a grid-stride copy whose CTAs live as long as the copy: two blocks per SM keep ~1 MB of PCIe
    // requests in flight (far beyond the link's bandwidth-delay product) and leave the SMs' other
    // warp slots to the step's kernels and to the copy of the opposite direction
    const int64_t per_sm = opt(OPT_COPY_BLOCKS) > 0 ? opt(OPT_COPY_BLOCKS) : 2;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec == 4)
        copy_cols_kernel<float4><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(d), reinterpret_cast<const float4*>(s),
                                                              rows, row_stride / 4, ncols / 4);
    else if (vec == 2)
        copy_cols_kernel<float2><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float2*>(d), reinterpret_cast<const float2*>(s),
                                                              rows, row_stride / 2, ncols / 2);
    else
        copy_cols_kernel<float><<<(int)blocks, 256, 0, st>>>(d, s, rows, row_stride, ncols);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}

// DMA-engine alternative to copy_cols (kept for measurement: scripts/experiments/stage_bw.py):
// cudaMemcpy2DAsync of the same column block; kind 1 = host -> device, 2 = device -> host.
extern "C" int hbv_b200_memcpy2d(float* dst, const float* src, int64_t rows, int64_t row_stride,
                                 int32_t col0, int32_t ncols, int32_t kind, void* stream) {
    using namespace hbv;
    if (!dst || !src || rows <= 0 || ncols <= 0 || (kind != 1 && kind != 2)) { set_error("memcpy2d: bad arguments"); return HBV_E_SHAPE; }
    cudaError_t e = cudaMemcpy2DAsync(dst + col0, (size_t)row_stride * 4, src + col0, (size_t)row_stride * 4,
                                      (size_t)ncols * 4, (size_t)rows,
                                      kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}
