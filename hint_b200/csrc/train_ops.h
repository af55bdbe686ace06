// Host interface of the training-edge kernels (train_ops.cu); the C ABI wrappers are in capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace hint {

cudaError_t train_add_noise(const float* x, float* out, long long n, float sigma, unsigned long long seed, unsigned long long offset,
                            cudaStream_t st);
size_t train_nll_workspace_bytes();
// loss3 = {0.5*mean_b |z_b|^2 - mean_b J_b, 0.5*mean |z|^2, mean J}, J = sum of the n_logdets [B] vectors (host array of device pointers)
cudaError_t train_nll_loss(const float* z, const float* const* logdets, int n_logdets, long long B, int d, float* loss3, void* workspace,
                           cudaStream_t st);
cudaError_t train_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                            const long long* sizes, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_clamp,
                            long long step, cudaStream_t st);

}  // namespace hint
