// Host-side launch interface of the tcgen05 training kernel (tc3_kernels.cuh), compiled in its own translation unit.
#pragma once
#include <cuda_runtime.h>

#include "plan_tc3.h"

namespace hint {

struct T3Prog;

struct DevTc3 {
    T3Epi* epis = nullptr;
    T3Chunk* chunks = nullptr;
    int16_t* tab16 = nullptr;
    int* pack_src = nullptr;
    int* part_dst = nullptr;     // partial index -> parameter index (-1: unused slot)
    T3Prog* prog = nullptr;      // host copy of the kernel-parameter program
    int num_sms = 0;
};

cudaError_t tc3_setup(const T3Plan& t, int num_sms, DevTc3& d);
void tc3_free(DevTc3& d);
cudaError_t tc3_pack(const T3Plan& t, const DevTc3& d, const float* params, float* packed, cudaStream_t st);
int tc3_bwd_ctas(const DevTc3& d, long long B);
// fused backward + deterministic reduction of the per-CTA partial gradients into dparams
cudaError_t tc3_launch_bwd(const T3Plan& t, const DevTc3& d, int grid, const float* z, const float* cond, const float* packed,
                           const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc, float* partials,
                           float* dparams, long long B, cudaStream_t st, long long* prof = nullptr, float nll_scale = 0.f);

// forward / inverse transport with a T3K_FORWARD / T3K_INVERSE plan: z [B,d], logdet [B]
cudaError_t tc3_launch_transport(const T3Plan& t, const DevTc3& d, const float* x, const float* cond, const float* packed, float* z,
                                 float* logdet, long long B, cudaStream_t st);

// developer aid: one tile, stop after n_epi_limit epilogue steps, dump TMEM [128][512] + raw shared memory (floats)
cudaError_t tc3_debug_run(const T3Plan& t, const DevTc3& d, int n_epi_limit, const float* z, const float* cond, const float* packed,
                          const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc, float* partials, long long B,
                          float* dump, cudaStream_t st);

}  // namespace hint
