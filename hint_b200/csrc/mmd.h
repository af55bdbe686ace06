// Multi-kernel MMD of two sample sets (rejection_sampling.py:56-73), fused; see mmd.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace hint {
size_t mmd_workspace_bytes(long long n);
cudaError_t mmd_multi(const float* x, const float* y, long long n, int d, const float* widths, const float* exponents, int n_kernels,
                      float* out, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace hint
