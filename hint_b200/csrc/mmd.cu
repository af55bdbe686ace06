// The evaluation metric of the reference's sampling scripts (rejection_sampling.py:56-73 `multi_mmd`):
//     MMD^2(x, y) = mean_ij [ k(|x_i - x_j|^2) + k(|y_i - y_j|^2) - 2 k(|x_i - y_j|^2) ],   k(D) = sum_w C_w^a_w ((C_w + D) / a_w)^(-a_w)
// for two sets of n samples (the reference uses n = 4000).  The reference materialises three n x n Gram matrices, three distance
// matrices and three kernel matrices (nine 64 MB tensors at n = 4000, ~40 launches); here ONE kernel walks 64 x 64 pair tiles of the
// three pair sets with the two row blocks in shared memory, evaluates distance and kernels in registers (4 x 4 pairs per thread)
// and reduces to one fp64 partial per tile; a second single-CTA kernel sums the partials in a fixed order (deterministic).
// Distances are computed as sum_k (a_k - b_k)^2 (never negative; the reference's rx + ry - 2 x.y form needs its clamp at 0).
#include <cuda_runtime.h>

#include <cstdint>

#include "launch_count.h"
#include "mmd.h"

namespace hint {

namespace {

constexpr int kMmdTile = 64, kMmdThreads = 256, kMmdMaxKernels = 8, kMmdChunk = 32;

struct MmdArgs {
    const float* x; const float* y; long long n; int d; int nk;
    float scale[kMmdMaxKernels], a[kMmdMaxKernels], C[kMmdMaxKernels];   // scale = C^a * a^a : k(D) = scale * (C + D)^(-a)
    double* partial; int tiles;   // tiles per side
};

__device__ __forceinline__ float mmd_kernel_sum(const MmdArgs& g, float D) {
    float k = 0.f;
    for (int w = 0; w < g.nk; ++w) {
        const float b = g.C[w] + D;
        float p;
        if (g.a[w] == 1.f) p = 1.f / b;
        else if (g.a[w] == 0.5f) p = rsqrtf(b);
        else p = powf(b, -g.a[w]);
        k = fmaf(g.scale[w], p, k);
    }
    return k;
}

// blockIdx.z: 0 = (x, x) weight +1, 1 = (y, y) weight +1, 2 = (x, y) weight -2
__global__ void __launch_bounds__(kMmdThreads) mmd_tile_kernel(MmdArgs g) {
    __shared__ float As[kMmdTile][kMmdChunk + 1];
    __shared__ float Bs[kMmdTile][kMmdChunk + 1];
    __shared__ double red[kMmdThreads / 32];
    const float* A = blockIdx.z == 1 ? g.y : g.x;
    const float* Bm = blockIdx.z == 0 ? g.x : g.y;
    const long long i0 = (long long)blockIdx.y * kMmdTile, j0 = (long long)blockIdx.x * kMmdTile;
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;      // rows ti + 16 p, columns tj + 16 q
    float D[4][4] = {};
    for (int k0 = 0; k0 < g.d; k0 += kMmdChunk) {
        const int kc = g.d - k0 < kMmdChunk ? g.d - k0 : kMmdChunk;
        __syncthreads();
        for (int i = threadIdx.x; i < kMmdTile * kMmdChunk; i += kMmdThreads) {
            const int r = i / kMmdChunk, k = i - r * kMmdChunk;
            As[r][k] = (k < kc && i0 + r < g.n) ? __ldg(A + (i0 + r) * g.d + k0 + k) : 0.f;
            Bs[r][k] = (k < kc && j0 + r < g.n) ? __ldg(Bm + (j0 + r) * g.d + k0 + k) : 0.f;
        }
        __syncthreads();
        for (int k = 0; k < kc; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) { a[p] = As[ti + 16 * p][k]; b[p] = Bs[tj + 16 * p][k]; }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float t = a[p] - b[q]; D[p][q] = fmaf(t, t, D[p][q]); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (i0 + ti + 16 * p < g.n && j0 + tj + 16 * q < g.n) s += mmd_kernel_sum(g, D[p][q]);
    double sd = (double)s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kMmdThreads / 32; ++w) t += red[w];
        g.partial[((size_t)blockIdx.z * g.tiles + blockIdx.y) * g.tiles + blockIdx.x] = blockIdx.z == 2 ? -2.0 * t : t;
    }
}

__global__ void __launch_bounds__(kMmdThreads) mmd_finish_kernel(const double* __restrict__ partial, long long count, double inv_n2, float* out) {
    __shared__ double red[kMmdThreads];
    double t = 0.0;
    for (long long i = threadIdx.x; i < count; i += kMmdThreads) t += partial[i];      // fixed assignment and order
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = kMmdThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(red[0] * inv_n2);
}

}  // namespace

size_t mmd_workspace_bytes(long long n) {
    const long long tiles = (n + kMmdTile - 1) / kMmdTile;
    return sizeof(double) * 3 * (size_t)tiles * (size_t)tiles + 64;
}

cudaError_t mmd_multi(const float* x, const float* y, long long n, int d, const float* widths, const float* exponents, int n_kernels,
                      float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n < 1 || d < 1 || n_kernels < 1 || n_kernels > kMmdMaxKernels || n > 65535LL * kMmdTile) return cudaErrorInvalidValue;
    if (ws_bytes < mmd_workspace_bytes(n)) return cudaErrorInvalidValue;
    MmdArgs g;
    g.x = x; g.y = y; g.n = n; g.d = d; g.nk = n_kernels;
    for (int w = 0; w < n_kernels; ++w) {
        if (!(widths[w] > 0.f) || !(exponents[w] > 0.f)) return cudaErrorInvalidValue;
        g.C[w] = widths[w]; g.a[w] = exponents[w];
        g.scale[w] = powf(widths[w], exponents[w]) * powf(exponents[w], exponents[w]);     // C^a ((C + D) / a)^-a = C^a a^a (C + D)^-a
    }
    g.partial = static_cast<double*>(ws);
    g.tiles = (int)((n + kMmdTile - 1) / kMmdTile);
    mmd_tile_kernel<<<dim3(g.tiles, g.tiles, 3), kMmdThreads, 0, st>>>(g); HINT_LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    mmd_finish_kernel<<<1, kMmdThreads, 0, st>>>(g.partial, 3LL * g.tiles * g.tiles, 1.0 / ((double)n * (double)n), out); HINT_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace hint
