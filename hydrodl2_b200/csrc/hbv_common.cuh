// hbv_common.cuh — kernel-side descriptor, parameter addressing and error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hbv_b200.h"
#include "hbv_step.cuh"

namespace hbv {

// Device-side copy of hbv_desc_t with derived quantities (passed by value as a kernel argument).
struct KDesc {
    int T, B, nmul, n_par;
    int nvar, i_prcp, i_tmean, i_pet;
    int dyn_ncol, sta_ncol;
    int apply_sigmoid;
    int K;               // checkpoint interval
    int muwts_t_stride;
    int BPB;             // basins per CTA
    float nearzero, dt, inv_dt;
    int src[HBV_MAX_PAR];
    int col[HBV_MAX_PAR];
    float lo[HBV_MAX_PAR];
    float span[HBV_MAX_PAR];  // hi - lo
};

struct FwdPtrs {
    const float* forcing; const float* dyn; const float* sta; const uint8_t* drop;
    const float* attrs; const float* muwts; const float* state_in;
    float* state_out; float* flux[HBV_MAX_FLUX]; float* state_series; float* ckpt;
};

struct BwdPtrs {
    const float* forcing; const float* dyn; const float* sta; const uint8_t* drop;
    const float* attrs; const float* muwts; const float* ckpt;
    const float* gflux[HBV_MAX_FLUX]; const float* gstate_out; const float* gstate_series;
    float* gdyn; float* gsta; float* gstate_in;
};

// value in [0,1] (or raw) -> physical parameter
__device__ __forceinline__ float descale(const KDesc& d, int i, float raw) {
    const float v = d.apply_sigmoid ? sigmoidf_(raw) : raw;
    return v * d.span[i] + d.lo[i];
}

// d(par)/d(raw) given the physical value is not needed: recompute from raw.
__device__ __forceinline__ float descale_grad(const KDesc& d, int i, float raw) {
    if (d.apply_sigmoid) {
        const float s = sigmoidf_(raw);
        return d.span[i] * s * (1.0f - s);
    }
    return d.span[i];
}

void set_error(const char* msg);
void count_launch(int n = 1);

}  // namespace hbv
