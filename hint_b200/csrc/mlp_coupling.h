// Fused FrEIA-style affine couplings with four-layer fully connected subnets; see mlp_coupling.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace hint {

// params / dparams: 16 device pointers, [s-net, t-net] x [fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc2b.weight, fc2b.bias, fc3.weight,
// fc3.bias], weights in PyTorch's [out][in] row-major layout.
bool mc_supported(int du, int dv, int H);
cudaError_t mc_forward(const float* u, int du, const float* v, int dv, int H, const float* const* params, float clamp, int rev, long long B,
                       float* y, float* logdet, cudaStream_t st);
size_t mc_workspace_bytes(int du, int dv, int H, long long B);
cudaError_t mc_backward(const float* u, int du, const float* v, int dv, int H, const float* const* params, float clamp, long long B,
                        const float* dy, const float* dlogdet, float* du_grad, float* dv_grad, float* const* dparams, void* ws,
                        size_t ws_bytes, cudaStream_t st);

// out[N][K] = Dm^T X (Dm [B][N], X [B][K]), fp32-grade, deterministic
size_t mc_xt_y_workspace_bytes(int N, int K);
cudaError_t mc_xt_y(const float* Dm, const float* X, int N, int K, long long B, float* out, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace hint
