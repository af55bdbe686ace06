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
    MmaParamProg* host_prog = nullptr;   // [3]: forward, inverse, backward programs in kernel-parameter form (host memory)
    int* pack_src = nullptr;
    int* unpack_src = nullptr;
};

// The packed operands are written `mma_weight_copies()` times (stride mma_copy_stride floats): CTA b reads copy b % copies.
// All CTAs walk the same op stream nearly in lock-step, so one copy makes 2*SMs CTAs hit the same few L2 lines at once.
constexpr int kMmaWeightCopies = 1;   // measured on B200 (d=43 hint_8): 1, 4, 16, 64 copies within 5% - L2 is not the limiter
int mma_weight_copies();
long long mma_copy_stride(const MmaPlan& m);

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
