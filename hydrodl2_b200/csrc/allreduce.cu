// allreduce.cu — one-shot all-reduce (sum) of a small vector over NVLink peer memory.
//
// The only collective of a training step on this path is the sum of the shared-parameter gradient
// (a few hundred floats).  Through NCCL that is an eager call after each
// graph-replayed step, ~30 us at 8 GPUs on a 465 us step, not capturable without side effects
// (DESIGN.md §5).  Here every rank owns a symmetric comm buffer whose peer addresses all ranks
// know (torch.distributed._symmetric_memory.rendezvous -> buffer_ptrs_dev); one CTA per rank
//   1. stores its vector into slot [parity][rank] of EVERY rank's buffer (plain stores to peer
//      memory: NVLink / NVSwitch), fences, and raises flag [parity][rank] on every rank
//      (st.release.sys);
//   2. waits until all `world` flags of its OWN buffer carry this step's number (ld.acquire.sys);
//   3. sums the `world` slots of its own buffer in rank order — the same order on every rank, so
//      all ranks get bit-identical results.
// The step number lives in the buffer and is advanced by the kernel itself, so the launch is a
// plain kernel node: it is captured in the CUDA graph of the step and replayed.  Slots are
// double-buffered by step parity: a rank can only write step s+2 after passing the barrier of
// step s+1, which needs every other rank's s+1 flag, which a rank raises only after its step-s
// kernel (the reader of the slot) has finished.
// Buffer layout (floats): data [2][world][ncap] | flags u32 [2][world] | step u32 | error u32.
#include "hbv_common.cuh"

namespace hbv {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
oneshot_allreduce_kernel(float* const* __restrict__ peers, int rank, int world, const float* __restrict__ in,
                         float* __restrict__ out, int n, int ncap, long long timeout_clocks) {
    float* const mine = peers[rank];
    uint32_t* const my_flags = reinterpret_cast<uint32_t*>(mine + (size_t)2 * world * ncap);
    uint32_t* const my_step = my_flags + 2 * world;
    __shared__ uint32_t s_step;
    if (threadIdx.x == 0) s_step = *my_step + 1;
    __syncthreads();
    const uint32_t step = s_step;
    const int par = (int)(step & 1u);
    for (int q = 0; q < world; ++q) {
        float* dst = peers[q] + ((size_t)par * world + rank) * ncap;
        for (int c = threadIdx.x; c < n; c += blockDim.x) dst[c] = in[c];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        uint32_t* f = reinterpret_cast<uint32_t*>(peers[threadIdx.x] + (size_t)2 * world * ncap) + par * world + rank;
        st_release_sys(f, step);
        const uint32_t* w = my_flags + par * world + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(w) - step) < 0) {
            if (clock64() - t0 > timeout_clocks) { my_step[1] = 1u; break; }      // a peer never arrived
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < world; ++q) a += mine[((size_t)par * world + q) * ncap + c];
        out[c] = a;
    }
    if (threadIdx.x == 0) *my_step = step;
}

}  // namespace hbv

extern "C" int64_t hbv_b200_allreduce_buffer_floats(int32_t world, int32_t n) {
    if (world <= 0 || n <= 0) return -1;
    const int64_t ncap = ((int64_t)n + 31) / 32 * 32;
    return 2 * world * ncap + 2 * world + 2;
}

extern "C" int hbv_b200_oneshot_allreduce(float* const* peer_bufs_dev, int32_t rank, int32_t world,
                                          const float* in, float* out, int32_t n, void* stream) {
    using namespace hbv;
    if (!peer_bufs_dev || !in || !out || world <= 0 || world > 64 || rank < 0 || rank >= world || n <= 0) {
        set_error("oneshot_allreduce: bad arguments");
        return HBV_E_SHAPE;
    }
    const int ncap = (n + 31) / 32 * 32;
    const long long timeout = 40LL * 1000 * 1000 * 1000;       // ~20 s of SM clocks: a peer that never arrives
    oneshot_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(peer_bufs_dev, rank, world, in, out, n, ncap, timeout);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error(cudaGetErrorString(e));
    return (int)e;
}
