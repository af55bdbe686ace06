// Host-side launch interface of the warp-MMA kernels (mma_kernels.cuh), compiled in its own translation unit
// (mma_launch.cu) so the C-ABI glue and the other kernel families build in parallel with it.
#pragma once
#include <cuda_runtime.h>

#include "plan_mma.h"

namespace hint {

struct DevMmaSchedule {
    WOp* prog = nullptr;
    Ep* eps = nullptr;
    int max_ctas = 0;     // SMs x occupancy (min over the TF32 and 3xTF32 instantiations)
};

struct DevMma {
    DevMmaSchedule fwd, bwd;
    int* pack_src = nullptr;
    int* unpack_src = nullptr;
};

cudaError_t mma_setup(const MmaPlan& m, int num_sms, DevMma& d);
void mma_free(DevMma& d);

// hi/lo: packed operand buffers of m.n_packed floats each (lo used only when x3)
cudaError_t mma_pack(const MmaPlan& m, const DevMma& d, const float* params, float* hi, float* lo, cudaStream_t st);
cudaError_t mma_launch_fwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, const float* x, const float* c,
                           const float* hi, const float* lo, float* z, float* logdet, long long B, int rev, cudaStream_t st);
// grid = number of CTAs = number of partial-gradient buffers in `partials`
cudaError_t mma_launch_bwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, int grid, const float* z, const float* c,
                           const float* hi, const float* lo, const float* dz, const float* dlogdet, float* x_rec, float* dx,
                           float* dc, float* partials, long long B, cudaStream_t st);

}  // namespace hint
