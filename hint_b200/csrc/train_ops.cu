// The training edge around the coupling blocks, as three small streaming kernels so that a training step
// launches nothing but this library's kernels (+ NCCL):
//   hint_add_noise   x' = x + sigma * N(0,1)                  train_unconditional.py:121-123 (`x += noise * randn`)
//   hint_nll_loss    0.5 * sum(z^2)/B - sum(logdet)/B          train_unconditional.py:128-132
//   hint_adam_step   clamp(grad, +-c) then Adam with L2 decay  train_unconditional.py:141-144, 174-176 (torch.optim.Adam)
// The gradient of the loss is never materialised: hint_backward_nll generates dz = z/B, dlogdet = -1/B in the tile load.
// All three are HBM-bound; 128-bit accesses, grid = a multiple of the SM count.
#include <cuda_runtime.h>

#include <cstdint>

#include "launch_count.h"
#include "train_ops.h"

namespace hint {

namespace {

// ---- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") ----------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// two uniforms -> two standard normals (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
    const float v = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    n0 = r * c; n1 = r * s;
}

__global__ void __launch_bounds__(256) hint_add_noise_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, float sigma,
                                                             unsigned long long seed, unsigned long long offset) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float g0, g1, g2, g3;
        box_muller(c[0], c[1], g0, g1);
        box_muller(c[2], c[3], g2, g3);
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        v.x = fmaf(sigma, g0, v.x); v.y = fmaf(sigma, g1, v.y); v.z = fmaf(sigma, g2, v.z); v.w = fmaf(sigma, g3, v.w);
        reinterpret_cast<float4*>(out)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // tail (n not a multiple of 4)
        const long long i = (n4 << 2) + threadIdx.x;
        uint32_t c[4] = {(uint32_t)n4, (uint32_t)(n4 >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float g[4];
        box_muller(c[0], c[1], g[0], g[1]);
        box_muller(c[2], c[3], g[2], g[3]);
        out[i] = fmaf(sigma, g[threadIdx.x], x[i]);
    }
}

// ---- NLL: two-stage deterministic reduction ----------------------------------------------------------------------------
constexpr int kNllBlocks = 148 * 4, kNllThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

constexpr int kNllMaxJ = 64;
struct NllJ { const float* p[kNllMaxJ]; int n; };

__global__ void __launch_bounds__(kNllThreads) hint_nll_partial_kernel(const float* __restrict__ z, const __grid_constant__ NllJ Js, long long nz,
                                                                       long long B, double* __restrict__ partial) {
    __shared__ double sh[kNllThreads / 32];
    double a = 0.0, b = 0.0;
    const long long n4 = nz >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(z) + i);
        a += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (nz & 3)) { const float v = z[(n4 << 2) + threadIdx.x]; a += (double)(v * v); }
    for (int q = 0; q < Js.n; ++q) {   // the per-block log-determinants are summed here: no elementwise add kernels on the step
        const float* __restrict__ J = Js.p[q];
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) b += (double)__ldg(J + i);
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
}
__global__ void __launch_bounds__(kNllThreads) hint_nll_final_kernel(const double* __restrict__ partial, int nblocks, long long B, float* __restrict__ loss) {
    __shared__ double sh[kNllThreads / 32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { loss[0] = (float)((0.5 * a - b) / (double)B); loss[1] = (float)(0.5 * a / (double)B); loss[2] = (float)(b / (double)B); }
}

// ---- clamp + Adam over up to kAdamMaxTensors flat tensors per launch -------------------------------------------------
constexpr int kAdamMaxTensors = 48;
struct AdamArgs {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    long long n[kAdamMaxTensors];
    int count;
    float lr, b1, b2, eps, wd, clamp, bc1, bc2s;   // bc1 = 1 - b1^t, bc2s = sqrt(1 - b2^t)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& A) {
    if (A.clamp > 0.f && g == g) g = fminf(fmaxf(g, -A.clamp), A.clamp);   // train_unconditional.py:141-142 (NaN propagates as in torch.clamp_)
    g = fmaf(A.wd, p, g);                                          // torch.optim.Adam: L2 decay folded into the gradient
    m = fmaf(A.b1, m, (1.f - A.b1) * g);
    v = fmaf(A.b2, v, (1.f - A.b2) * g * g);
    const float denom = sqrtf(v) / A.bc2s + A.eps;
    p -= (A.lr / A.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) hint_adam_kernel(const __grid_constant__ AdamArgs A) {
    const int t = blockIdx.y;
    if (t >= A.count) return;
    float* p = A.p[t]; const float* g = A.g[t]; float* m = A.m[t]; float* v = A.v[t];
    const long long n = A.n[t], n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam_one(pp.x, gg.x, mm.x, vv.x, A); adam_one(pp.y, gg.y, mm.y, vv.y, A);
        adam_one(pp.z, gg.z, mm.z, vv.z, A); adam_one(pp.w, gg.w, mm.w, vv.w, A);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        adam_one(p[i], g[i], m[i], v[i], A);
    }
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

cudaError_t train_add_noise(const float* x, float* out, long long n, float sigma, unsigned long long seed, unsigned long long offset,
                            cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (!al16(x) || !al16(out)) return cudaErrorMisalignedAddress;
    const long long want = ((n >> 2) + 255) / 256;
    const int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    hint_add_noise_kernel<<<blocks, 256, 0, st>>>(x, out, n, sigma, seed, offset); HINT_LAUNCHED();
    return cudaGetLastError();
}

size_t train_nll_workspace_bytes() { return sizeof(double) * 2 * kNllBlocks; }

cudaError_t train_nll_loss(const float* z, const float* const* logdets, int n_logdets, long long B, int d, float* loss3, void* workspace,
                           cudaStream_t st) {
    if (B <= 0 || n_logdets < 0 || n_logdets > kNllMaxJ) return cudaErrorInvalidValue;
    if (!al16(z) || !al16(workspace)) return cudaErrorMisalignedAddress;
    double* partial = reinterpret_cast<double*>(workspace);
    NllJ Js{};
    Js.n = n_logdets;
    for (int i = 0; i < n_logdets; ++i) { if (!logdets[i]) return cudaErrorInvalidValue; Js.p[i] = logdets[i]; }
    hint_nll_partial_kernel<<<kNllBlocks, kNllThreads, 0, st>>>(z, Js, B * (long long)d, B, partial); HINT_LAUNCHED();
    hint_nll_final_kernel<<<1, kNllThreads, 0, st>>>(partial, kNllBlocks, B, loss3); HINT_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t train_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                            const long long* sizes, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_clamp,
                            long long step, cudaStream_t st) {
    if (n_tensors <= 0) return cudaSuccess;
    if (step < 1) return cudaErrorInvalidValue;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    for (int t0 = 0; t0 < n_tensors; t0 += kAdamMaxTensors) {
        AdamArgs A{};
        A.count = n_tensors - t0 < kAdamMaxTensors ? n_tensors - t0 : kAdamMaxTensors;
        long long nmax = 0;
        for (int i = 0; i < A.count; ++i) {
            A.p[i] = params[t0 + i]; A.g[i] = grads[t0 + i]; A.m[i] = exp_avg[t0 + i]; A.v[i] = exp_avg_sq[t0 + i]; A.n[i] = sizes[t0 + i];
            if (!A.p[i] || !A.g[i] || !A.m[i] || !A.v[i] || A.n[i] < 0) return cudaErrorInvalidValue;
            if (!al16(A.p[i]) || !al16(A.g[i]) || !al16(A.m[i]) || !al16(A.v[i])) return cudaErrorMisalignedAddress;
            nmax = A.n[i] > nmax ? A.n[i] : nmax;
        }
        A.lr = lr; A.b1 = beta1; A.b2 = beta2; A.eps = eps; A.wd = weight_decay; A.clamp = grad_clamp;
        A.bc1 = (float)bc1; A.bc2s = (float)sqrt(bc2);
        const long long want = ((nmax >> 2) + 255) / 256;
        const int bx = (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
        hint_adam_kernel<<<dim3(bx, A.count), 256, 0, st>>>(A); HINT_LAUNCHED();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace hint
