// Host interface of the inter-block Householder mixing kernels (householder.cu); the C ABI wrappers are in capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace hint {

int hh_max_d();
cudaError_t hh_matrix(const float* Vs, int n, int d, float* W, cudaStream_t st);
cudaError_t hh_matrix_backward(const float* Vs, const float* W, const float* dW, int n, int d, float* dVs, cudaStream_t st);
cudaError_t hh_apply(const float* x, const float* W, long long B, int d, int transpose, float* y, cudaStream_t st);
size_t hh_wgrad_workspace_bytes(int d);
cudaError_t hh_wgrad(const float* x, const float* dz, long long B, int d, float* dW, void* workspace, cudaStream_t st);

}  // namespace hint
