// Host-side launch interface of the register-chained warp-MMA kernels (chain_kernels.cuh), compiled in its own
// translation unit (chain_launch.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "plan_chain.h"

namespace hint {

struct DevChain {
    int* pack_src = nullptr;
    int* unpack_src = nullptr;
    int num_sms = 0;
    size_t fwd_smem = 0, bwd_smem = 0;
    int bwd_ctas = 0;      // resident CTAs of the backward kernel (SMs x occupancy)
    int bwd_mt = 1, bwd_nw = 4;
};

// host-side envelope check: the tile state of the forward (L1 configuration) and backward kernels must fit shared memory
bool chain_fits(const Plan& p, const ChainPlan& c, std::string* why);
cudaError_t chain_setup(const Plan& p, const ChainPlan& c, int num_sms, DevChain& d);
void chain_free(DevChain& d);
cudaError_t chain_pack(const ChainPlan& c, const DevChain& d, const float* params, float* packed, cudaStream_t st);
// number of CTAs (= partial-gradient buffers) the backward launch uses for a batch of B samples
int chain_bwd_ctas(const ChainPlan& c, const DevChain& d, long long B);
cudaError_t chain_launch_bwd(const Plan& p, const ChainPlan& c, const DevChain& d, int grid, const float* z, const float* cond,
                             const float* packed, const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc,
                             float* partials, long long B, cudaStream_t st, float nll_scale = 0.f);
cudaError_t chain_launch_fwd(const Plan& p, const ChainPlan& c, const DevChain& d, const float* x, const float* cond,
                             const float* packed, float* z, float* logdet, long long B, int rev, cudaStream_t st);

}  // namespace hint
