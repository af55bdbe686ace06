// Inter-block orthogonal mixing: every reference model is  [HINT] (-> x W -> HINT) x (n_blocks - 1)  with
// W = prod_i (I - 2 v_i v_i^T / |v_i|^2) a product of n_reflections = d Householder reflections (FrEIA `HouseholderPerm`,
// configs/uci_data/miniboone_hint_8.py:60-63 `fixed: True`; configs/plus_shape/unconditional_hint_4_3.py:60-71 `fixed: False`,
// trainable).  FrEIA's source is not part of the reference: the definition is the published one, parity-unpinned.
//
//   hh_matrix_kernel       W from Vs: one CTA, W resident in shared memory, d reflections of O(d^2) each
//   hh_matrix_bwd_kernel   dVs from dW WITHOUT storing the partial products: reflections are involutions, so the chain is
//                          walked backwards by W_{k-1} = W_k H_k while dW is pulled back by G_{k-1} = G_k H_k; per reflection
//                          only matrix-vector products (a = G v, b = W'^T a, c = W' v, e = G^T c):
//                              dv = -2/s (b + e) + 4 (c . a) / s^2 v,   s = |v|^2,  W' = W_{k-1}
//   hh_apply_kernel        y = x W or x W^T: FP32 FFMA (the mixing must stay orthogonal to 1e-6: no TF32), W resident in shared
//                          memory per CTA, 4 x 4 register tiles, 128-row tiles; HBM-bound for d <= 43, FFMA-bound at d = 100
//   hh_wgrad_kernel        dW = x^T dz: per-CTA partials in registers over its row tiles, fixed-order second stage (deterministic)
#include <cuda_runtime.h>

#include <cstdint>

#include "householder.h"
#include "launch_count.h"

namespace hint {

namespace {

constexpr int kHhThreads = 256;
constexpr int kHhMaxD = 128;

__device__ __forceinline__ float hh_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < kHhThreads / 32; ++w) r += red[w];
    return r;
}

// u = M v (rows) or M^T v (cols) for a [d][dp] shared-memory matrix; result in out[0..d)
__device__ __forceinline__ void hh_matvec(const float* M, const float* v, float* out, int d, int dp, bool transpose) {
    for (int r = threadIdx.x; r < d; r += kHhThreads) {
        float a = 0.f;
        if (!transpose) for (int c = 0; c < d; ++c) a = fmaf(M[r * dp + c], v[c], a);
        else for (int c = 0; c < d; ++c) a = fmaf(M[c * dp + r], v[c], a);
        out[r] = a;
    }
    __syncthreads();
}
// M -= 2 u v^T / s
__device__ __forceinline__ void hh_rank1(float* M, const float* u, const float* v, float two_over_s, int d, int dp) {
    for (int i = threadIdx.x; i < d * d; i += kHhThreads) {
        const int r = i / d, c = i - r * d;
        M[r * dp + c] = fmaf(-two_over_s * u[r], v[c], M[r * dp + c]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kHhThreads) hh_matrix_kernel(const float* __restrict__ Vs, int n, int d, float* __restrict__ W) {
    extern __shared__ __align__(16) float sm[];
    const int dp = d + 1;
    float* Ws = sm;                 // [d][dp]
    float* v = Ws + d * dp;         // [d]
    float* u = v + kHhMaxD;         // [d]
    float* red = u + kHhMaxD;       // [8]
    for (int i = threadIdx.x; i < d * d; i += kHhThreads) { const int r = i / d, c = i - r * d; Ws[r * dp + c] = r == c ? 1.f : 0.f; }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        float part = 0.f;
        for (int c = threadIdx.x; c < d; c += kHhThreads) { const float t = Vs[(size_t)k * d + c]; v[c] = t; part = fmaf(t, t, part); }
        const float s = hh_block_sum(part, red);
        hh_matvec(Ws, v, u, d, dp, false);
        hh_rank1(Ws, u, v, 2.f / s, d, dp);
    }
    for (int i = threadIdx.x; i < d * d; i += kHhThreads) { const int r = i / d, c = i - r * d; W[i] = Ws[r * dp + c]; }
}

__global__ void __launch_bounds__(kHhThreads) hh_matrix_bwd_kernel(const float* __restrict__ Vs, const float* __restrict__ W,
                                                                   const float* __restrict__ dW, int n, int d, float* __restrict__ dVs) {
    extern __shared__ __align__(16) float sm[];
    const int dp = d + 1;
    float* Ws = sm;                   // W_k, walked back to W_{k-1}
    float* Gs = Ws + d * dp;          // G_k
    float* v = Gs + d * dp;
    float* a = v + kHhMaxD;
    float* b = a + kHhMaxD;
    float* c = b + kHhMaxD;
    float* e = c + kHhMaxD;
    float* red = e + kHhMaxD;
    for (int i = threadIdx.x; i < d * d; i += kHhThreads) {
        const int r = i / d, cc = i - r * d;
        Ws[r * dp + cc] = W[i];
        Gs[r * dp + cc] = dW[i];
    }
    __syncthreads();
    for (int k = n - 1; k >= 0; --k) {
        float part = 0.f;
        for (int q = threadIdx.x; q < d; q += kHhThreads) { const float t = Vs[(size_t)k * d + q]; v[q] = t; part = fmaf(t, t, part); }
        const float s = hh_block_sum(part, red);
        // W_{k-1} = W_k H_k  (H_k is its own inverse)
        hh_matvec(Ws, v, c, d, dp, false);          // c = W_k v  (temporarily)
        hh_rank1(Ws, c, v, 2.f / s, d, dp);
        hh_matvec(Gs, v, a, d, dp, false);          // a = G_k v
        hh_matvec(Ws, a, b, d, dp, true);           // b = W_{k-1}^T a = M v
        hh_matvec(Ws, v, c, d, dp, false);          // c = W_{k-1} v
        hh_matvec(Gs, c, e, d, dp, true);           // e = G_k^T c = M^T v
        float pd = 0.f;
        for (int q = threadIdx.x; q < d; q += kHhThreads) pd = fmaf(c[q], a[q], pd);
        const float vMv = hh_block_sum(pd, red);
        for (int q = threadIdx.x; q < d; q += kHhThreads)
            dVs[(size_t)k * d + q] = -2.f / s * (b[q] + e[q]) + 4.f * vMv / (s * s) * v[q];
        hh_rank1(Gs, a, v, 2.f / s, d, dp);         // G_{k-1} = G_k H_k
    }
}

// y[r, :] = x[r, :] W (or W^T).  CTA = 128 rows; thread (rg = tid / 8, cg = tid % 8) owns rows 4 rg .. 4 rg + 3 and the column
// quads cg, cg + 8, ...  Shared memory: W [d][dq4] (dq4 = d rounded up to 4, zero padded), x tile [128][d + 1].
constexpr int kHhRows = 128;

__global__ void __launch_bounds__(kHhThreads) hh_apply_kernel(const float* __restrict__ x, const float* __restrict__ W, long long B, int d,
                                                              int transpose, float* __restrict__ y) {
    extern __shared__ __align__(16) float sm[];
    const int d4 = (d + 3) & ~3, xp = d + 1;
    float* Ws = sm;                    // [d][d4]
    float* Xs = Ws + d * d4;           // [128][xp]
    for (int i = threadIdx.x; i < d * d4; i += kHhThreads) {
        const int k = i / d4, j = i - k * d4;
        Ws[i] = j < d ? (transpose ? W[(size_t)j * d + k] : W[(size_t)k * d + j]) : 0.f;
    }
    const int rg = threadIdx.x >> 3, cg = threadIdx.x & 7;
    const long long ntiles = (B + kHhRows - 1) / kHhRows;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * kHhRows;
        const int rows = (int)((B - row0) < kHhRows ? (B - row0) : kHhRows);
        __syncthreads();
        const float* gx = x + row0 * d;
        for (int i = threadIdx.x; i < kHhRows * d; i += kHhThreads) {
            const int r = i / d, k = i - r * d;
            Xs[r * xp + k] = r < rows ? __ldg(gx + i) : 0.f;
        }
        __syncthreads();
        for (int q = cg; q * 4 < d; q += 8) {
            float acc[4][4] = {};
            const float* xr = Xs + (4 * rg) * xp;
            for (int k = 0; k < d; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + k * d4 + 4 * q);
                const float x0 = xr[k], x1 = xr[xp + k], x2 = xr[2 * xp + k], x3 = xr[3 * xp + k];
                acc[0][0] = fmaf(x0, w.x, acc[0][0]); acc[0][1] = fmaf(x0, w.y, acc[0][1]); acc[0][2] = fmaf(x0, w.z, acc[0][2]); acc[0][3] = fmaf(x0, w.w, acc[0][3]);
                acc[1][0] = fmaf(x1, w.x, acc[1][0]); acc[1][1] = fmaf(x1, w.y, acc[1][1]); acc[1][2] = fmaf(x1, w.z, acc[1][2]); acc[1][3] = fmaf(x1, w.w, acc[1][3]);
                acc[2][0] = fmaf(x2, w.x, acc[2][0]); acc[2][1] = fmaf(x2, w.y, acc[2][1]); acc[2][2] = fmaf(x2, w.z, acc[2][2]); acc[2][3] = fmaf(x2, w.w, acc[2][3]);
                acc[3][0] = fmaf(x3, w.x, acc[3][0]); acc[3][1] = fmaf(x3, w.y, acc[3][1]); acc[3][2] = fmaf(x3, w.z, acc[3][2]); acc[3][3] = fmaf(x3, w.w, acc[3][3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 4 * rg + i;
                if (r >= rows) continue;
                float* gy = y + (row0 + r) * d + 4 * q;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * q + j < d) gy[j] = acc[i][j];
            }
        }
    }
}

// partial[cta][i][j] = sum over the CTA's rows of x[r][i] dz[r][j];  thread (ig = tid / 16, jg = tid % 16) owns the 4 x 4 blocks
// (ig + 16 a, jg + 16 b), a, b < 2  -> d <= 128.
__global__ void __launch_bounds__(kHhThreads) hh_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, long long B, int d,
                                                              float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    const int dp = ((d + 3) & ~3) + 4;           // row pitch in floats (16-byte aligned rows, padded)
    constexpr int TR = 32;                        // rows per step
    float* Xs = sm;                               // [TR][dp]
    float* Zs = Xs + TR * dp;                     // [TR][dp]
    const int ig = threadIdx.x >> 4, jg = threadIdx.x & 15;
    float acc[2][2][4][4] = {};
    const long long nsteps = (B + TR - 1) / TR;
    for (long long st = blockIdx.x; st < nsteps; st += gridDim.x) {
        const long long row0 = st * TR;
        const int rows = (int)((B - row0) < TR ? (B - row0) : TR);
        __syncthreads();
        for (int i = threadIdx.x; i < TR * dp; i += kHhThreads) {
            const int r = i / dp, k = i - r * dp;
            const bool ok = r < rows && k < d;
            Xs[i] = ok ? __ldg(x + (row0 + r) * d + k) : 0.f;
            Zs[i] = ok ? __ldg(dz + (row0 + r) * d + k) : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < TR; ++r) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int i0 = 4 * (ig + 16 * a);
                if (i0 >= d) continue;
                const float4 xv = *reinterpret_cast<const float4*>(Xs + r * dp + i0);
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int j0 = 4 * (jg + 16 * b);
                    if (j0 >= d) continue;
                    const float4 zv = *reinterpret_cast<const float4*>(Zs + r * dp + j0);
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, zs[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[a][b][p][q] = fmaf(xs[p], zs[q], acc[a][b][p][q]);
                }
            }
        }
    }
    float* out = partial + (size_t)blockIdx.x * d * d;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = 4 * (ig + 16 * a) + p, j = 4 * (jg + 16 * b) + q;
                    if (i < d && j < d) out[i * d + j] = acc[a][b][p][q];
                }
}
__global__ void hh_wgrad_reduce_kernel(const float* __restrict__ partial, int nctas, int n, float* __restrict__ dW) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < nctas; ++q) a += partial[(size_t)q * n + i];
        dW[i] = a;
    }
}

int hh_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

}  // namespace

int hh_max_d() { return kHhMaxD; }

cudaError_t hh_matrix(const float* Vs, int n, int d, float* W, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || n < 0) return cudaErrorInvalidValue;
    const size_t smem = sizeof(float) * ((size_t)d * (d + 1) + 2 * kHhMaxD + 8);
    cudaError_t e = cudaFuncSetAttribute((const void*)hh_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hh_matrix_kernel<<<1, kHhThreads, smem, st>>>(Vs, n, d, W); HINT_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t hh_matrix_backward(const float* Vs, const float* W, const float* dW, int n, int d, float* dVs, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || n < 0) return cudaErrorInvalidValue;
    const size_t smem = sizeof(float) * (2 * (size_t)d * (d + 1) + 5 * kHhMaxD + 8);
    cudaError_t e = cudaFuncSetAttribute((const void*)hh_matrix_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hh_matrix_bwd_kernel<<<1, kHhThreads, smem, st>>>(Vs, W, dW, n, d, dVs); HINT_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t hh_apply(const float* x, const float* W, long long B, int d, int transpose, float* y, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || B < 0) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    const int d4 = (d + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)d * d4 + (size_t)kHhRows * (d + 1));
    cudaError_t e = cudaFuncSetAttribute((const void*)hh_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long ntiles = (B + kHhRows - 1) / kHhRows;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int grid = (int)(ntiles < (long long)hh_sms() * per_sm ? ntiles : (long long)hh_sms() * per_sm);
    hh_apply_kernel<<<grid, kHhThreads, smem, st>>>(x, W, B, d, transpose, y); HINT_LAUNCHED();
    return cudaGetLastError();
}

size_t hh_wgrad_workspace_bytes(int d) { return sizeof(float) * (size_t)hh_sms() * 2 * d * d; }

cudaError_t hh_wgrad(const float* x, const float* dz, long long B, int d, float* dW, void* workspace, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || B < 0) return cudaErrorInvalidValue;
    const int dp = ((d + 3) & ~3) + 4;
    const size_t smem = sizeof(float) * 2 * 32 * dp;
    const long long nsteps = (B + 31) / 32;
    int grid = (int)(nsteps < (long long)hh_sms() * 2 ? nsteps : (long long)hh_sms() * 2);
    if (grid < 1) grid = 1;
    float* partial = reinterpret_cast<float*>(workspace);
    hh_wgrad_kernel<<<grid, kHhThreads, smem, st>>>(x, dz, B, d, partial); HINT_LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = d * d;
    hh_wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, grid, n, dW); HINT_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace hint
