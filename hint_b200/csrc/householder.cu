// Inter-block orthogonal mixing: every reference model is  [HINT] (-> x W -> HINT) x (n_blocks - 1)  with
// W = prod_i (I - 2 v_i v_i^T / |v_i|^2) a product of n_reflections = d Householder reflections (FrEIA `HouseholderPerm`,
// configs/uci_data/miniboone_hint_8.py:60-63 `fixed: True`; configs/plus_shape/unconditional_hint_4_3.py:60-71 `fixed: False`,
// trainable).  FrEIA's source is not part of the reference: the definition is the published one, parity-unpinned.
//
//   hh_matrix_kernel       W from Vs: the chain of reflections is row-wise, a warp carries 4 rows of W in registers through it
//   hh_matrix_bwd_kernel   dVs from dW WITHOUT storing the partial products: reflections are involutions, so the chain is
//                          walked backwards by W_{k-1} = W_k H_k while dW is pulled back by G_{k-1} = G_k H_k; per reflection
//                          only matrix-vector products (a = G v, b = W'^T a, c = W' v, e = G^T c):
//                              dv = -2/s (b + e) + 4 (c . a) / s^2 v,   s = |v|^2,  W' = W_{k-1}
//   hh_apply_kernel        y = x W or x W^T: error-compensated 3 x TF32 warp MMAs (fp32-grade: the mixing must stay orthogonal to
//                          1e-6), W resident in shared memory, per-warp cp.async double-buffered row ranges
//   hh_wgrad               dW = x^T dz on the couplings' weight-gradient kernel (mlp_coupling.cu: 3 x TF32 MMAs, deterministic)
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "householder.h"
#include "mlp_coupling.h"
#include "launch_count.h"

namespace hint {

namespace {

constexpr int kHhThreads = 256;
constexpr int kHhMaxD = 128;

// The reflections act on W from the right, W_k = W_{k-1} (I - 2 v v^T / s): every ROW of W goes through the whole chain on its
// own, row <- row - (2 (row . v) / s) v.  A warp owns kHhRowsPerWarp rows (lane l holds elements l, l + 32, l + 64, l + 96 of each,
// d <= 128), the dot products are butterfly reductions and the rows' chains interleave: no barrier, ~60 cycles per reflection
// instead of three CTA-wide passes over a shared-memory matrix.
constexpr int kHhRowsPerWarp = 4;
constexpr int kHhLaneElems = kHhMaxD / 32;

__device__ __forceinline__ float hh_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void hh_load_v(const float* __restrict__ Vs, int k, int d, int lane, float (&v)[kHhLaneElems]) {
#pragma unroll
    for (int m = 0; m < kHhLaneElems; ++m) { const int j = lane + 32 * m; v[m] = j < d ? __ldg(Vs + (size_t)k * d + j) : 0.f; }
}
__device__ __forceinline__ float hh_dot(const float (&a)[kHhLaneElems], const float (&b)[kHhLaneElems]) {
    float p = 0.f;
#pragma unroll
    for (int m = 0; m < kHhLaneElems; ++m) p = fmaf(a[m], b[m], p);
    return hh_warp_sum(p);
}

__global__ void __launch_bounds__(128) hh_matrix_kernel(const float* __restrict__ Vs, int n, int d, float* __restrict__ W) {
    const int lane = threadIdx.x & 31, r0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kHhRowsPerWarp;
    if (r0 >= d) return;
    float row[kHhRowsPerWarp][kHhLaneElems], v[kHhLaneElems], vn[kHhLaneElems];
#pragma unroll
    for (int i = 0; i < kHhRowsPerWarp; ++i)
#pragma unroll
        for (int m = 0; m < kHhLaneElems; ++m) row[i][m] = (r0 + i == lane + 32 * m) ? 1.f : 0.f;
    if (n > 0) hh_load_v(Vs, 0, d, lane, vn);
    for (int k = 0; k < n; ++k) {
#pragma unroll
        for (int m = 0; m < kHhLaneElems; ++m) v[m] = vn[m];
        if (k + 1 < n) hh_load_v(Vs, k + 1, d, lane, vn);
        const float two_over_s = 2.f / hh_dot(v, v);
        float f[kHhRowsPerWarp];
#pragma unroll
        for (int i = 0; i < kHhRowsPerWarp; ++i) f[i] = two_over_s * hh_dot(row[i], v);
#pragma unroll
        for (int i = 0; i < kHhRowsPerWarp; ++i)
#pragma unroll
            for (int m = 0; m < kHhLaneElems; ++m) row[i][m] = fmaf(-f[i], v[m], row[i][m]);
    }
#pragma unroll
    for (int i = 0; i < kHhRowsPerWarp; ++i)
#pragma unroll
        for (int m = 0; m < kHhLaneElems; ++m) {
            const int j = lane + 32 * m;
            if (r0 + i < d && j < d) W[(size_t)(r0 + i) * d + j] = row[i][m];
        }
}

// Backward of the chain.  Walking back (k = n-1 .. 0) with W' = W_{k-1} = W_k H_k and G_{k-1} = G_k H_k is again row-wise; what
// couples the rows is only the OUTPUT dv_k = -2/s (W'^T a + G_k^T c) + 4 (c . a) / s^2 v  with a = G_k v, c = W' v = -W_k v:
// every warp adds its rows' terms into a shared buffer slot and the first d threads sum the slots in a fixed order
// (deterministic); the buffer is double-buffered, so ONE CTA barrier per reflection.  One CTA of 32 warps x 4 rows.
constexpr int kHhBwdWarps = kHhMaxD / kHhRowsPerWarp;

__global__ void __launch_bounds__(32 * kHhBwdWarps) hh_matrix_bwd_kernel(const float* __restrict__ Vs, const float* __restrict__ W,
                                                                         const float* __restrict__ dW, int n, int d, float* __restrict__ dVs) {
    __shared__ float P[2][kHhBwdWarps][kHhMaxD];
    __shared__ float Q[2][kHhBwdWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r0 = warp * kHhRowsPerWarp;
    const int nwarps = (d + kHhRowsPerWarp - 1) / kHhRowsPerWarp;
    float wr[kHhRowsPerWarp][kHhLaneElems], gr[kHhRowsPerWarp][kHhLaneElems], v[kHhLaneElems], vn[kHhLaneElems];
#pragma unroll
    for (int i = 0; i < kHhRowsPerWarp; ++i)
#pragma unroll
        for (int m = 0; m < kHhLaneElems; ++m) {
            const int j = lane + 32 * m;
            const bool in = r0 + i < d && j < d;
            wr[i][m] = in ? W[(size_t)(r0 + i) * d + j] : 0.f;
            gr[i][m] = in ? dW[(size_t)(r0 + i) * d + j] : 0.f;
        }
    if (n > 0) hh_load_v(Vs, n - 1, d, lane, vn);
    for (int k = n - 1; k >= 0; --k) {
        const int slot = k & 1;
#pragma unroll
        for (int m = 0; m < kHhLaneElems; ++m) v[m] = vn[m];
        if (k > 0) hh_load_v(Vs, k - 1, d, lane, vn);
        const float s = hh_dot(v, v), two_over_s = 2.f / s;
        if (warp < nwarps) {
            float c0[kHhRowsPerWarp], a[kHhRowsPerWarp];
#pragma unroll
            for (int i = 0; i < kHhRowsPerWarp; ++i) { c0[i] = hh_dot(wr[i], v); a[i] = hh_dot(gr[i], v); }
            float p[kHhLaneElems] = {}, q = 0.f;
#pragma unroll
            for (int i = 0; i < kHhRowsPerWarp; ++i) {
                const float c = -c0[i];                                // W_{k-1} v = W_k H_k v = -W_k v
                q = fmaf(c, a[i], q);
#pragma unroll
                for (int m = 0; m < kHhLaneElems; ++m) {
                    wr[i][m] = fmaf(-two_over_s * c0[i], v[m], wr[i][m]);   // row of W_{k-1}
                    p[m] = fmaf(a[i], wr[i][m], fmaf(c, gr[i][m], p[m]));   // a W'^T + c G_k^T terms
                    gr[i][m] = fmaf(-two_over_s * a[i], v[m], gr[i][m]);    // row of G_{k-1}
                }
            }
#pragma unroll
            for (int m = 0; m < kHhLaneElems; ++m) P[slot][warp][lane + 32 * m] = p[m];
            if (lane == 0) Q[slot][warp] = q;
        }
        __syncthreads();
        if ((int)threadIdx.x < d) {
            float bp = 0.f, vMv = 0.f;
            for (int w = 0; w < nwarps; ++w) { bp += P[slot][w][threadIdx.x]; vMv += Q[slot][w]; }
            const float vj = __ldg(Vs + (size_t)k * d + threadIdx.x);
            dVs[(size_t)k * d + threadIdx.x] = -two_over_s * bp + 4.f * vMv / (s * s) * vj;
        }
    }
}

// y[r, :] = x[r, :] W (or W^T) on the tensor cores with fp32-grade accuracy: both operands are split into tf32 hi + lo parts
// and every product is three `mma.m16n8k8.tf32` (lo*hi, hi*lo, hi*hi; the dropped lo*lo term is 2^-22 relative), fp32
// accumulation.  (An FFMA version of this kernel was bound by shared-memory wavefronts at 15 TFLOP/s: a 4 x 4 register tile needs
// one wavefront per two FFMA instructions, `profiles/ncu_r02_hh_apply.txt`.)  Warps are independent: a warp owns RW = 16 MT
// consecutive rows = one CONTIGUOUS range of x, fetched by 16-byte cp.async into its private double buffer (same layout as in
// global memory) while it multiplies the previous range; the products are staged in the consumed buffer and leave as 128-bit
// stores; no CTA barrier in the loop.  W is resident in shared memory, already split (hi and lo images, zero padded to [K8][NP],
// NP mod 16 = 8: B fragments conflict-free).
__device__ __forceinline__ void hh_cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void hh_cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t hh_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void hh_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = hh_tf32(v);
    lo = hh_tf32(v - __uint_as_float(hi));
}
__device__ __forceinline__ void hh_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// a contiguous range of n floats global -> shared (16-byte copies when the global side is aligned)
__device__ __forceinline__ void hh_range_fetch(float* dst, const float* src, int n, bool vec, int lane) {
    const int n4 = vec ? n >> 2 : 0;
    for (int i = lane; i < n4; i += 32) hh_cp_async16(dst + 4 * i, src + 4 * i);
    for (int i = 4 * n4 + lane; i < n; i += 32) hh_cp_async4(dst + i, src + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

constexpr int kHhApplyThreads = 512;

// MT m-tiles of 16 rows per warp, all NT <= NTC output n-tiles accumulated in one pass over K: (MT, NTC) = (2, 6) for d <= 48,
// (1, 16) for d <= 128.
template <int MT, int NTC>
__global__ void __launch_bounds__(kHhApplyThreads) hh_apply_kernel(const float* __restrict__ x, const float* __restrict__ W, long long B, int d,
                                                                   int transpose, float* __restrict__ y) {
    extern __shared__ __align__(16) float sm[];
    constexpr int RW = 16 * MT;
    const int K8 = (d + 7) & ~7, NT = K8 >> 3, NP = K8 + ((K8 & 15) == 8 ? 0 : 8), bufsz = (RW * d + 8 + 3) & ~3;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t* Wh = reinterpret_cast<uint32_t*>(sm);    // [K8][NP] tf32 hi parts, then [K8][NP] lo parts
    uint32_t* Wl = Wh + K8 * NP;
    float* Xw = sm + 2 * K8 * NP + warp * (2 * bufsz); // this warp's two row ranges, rows of pitch d (as in global memory) + 8 floats of slack
    for (int i = threadIdx.x; i < K8 * NP; i += blockDim.x) {
        const int k = i / NP, j = i - k * NP;
        hh_split((j < d && k < d) ? (transpose ? W[(size_t)j * d + k] : W[(size_t)k * d + j]) : 0.f, Wh[i], Wl[i]);
    }
    __syncthreads();
    const bool vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0, vec_out = (reinterpret_cast<uintptr_t>(y) & 15) == 0;   // range offsets are multiples of 64 bytes
    const long long nranges = (B + RW - 1) / RW, stride = (long long)gridDim.x * nw;
    auto floats_of = [&](long long r) { const long long left = B - r * RW; return (int)(left < RW ? left : RW) * d; };
    long long rng = (long long)blockIdx.x * nw + warp;
    if (rng < nranges) hh_range_fetch(Xw, x + rng * RW * d, floats_of(rng), vec_in, lane);
    int buf = 0;
    for (; rng < nranges; rng += stride, buf ^= 1) {
        float* Xb = Xw + buf * bufsz;
        const int n = floats_of(rng);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (rng + stride < nranges) hh_range_fetch(Xw + (buf ^ 1) * bufsz, x + (rng + stride) * RW * d, floats_of(rng + stride), vec_in, lane);
        float acc[MT][NTC][4] = {};
        for (int ks = 0; ks < NT; ++ks) {
            uint32_t ah[MT][4], al[MT][4];
            const bool k0 = 8 * ks + t < d, k1 = 8 * ks + t + 4 < d;   // columns beyond d hold the next row: mask them (last k-step only)
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const float* ap = Xb + (16 * i + g) * d + 8 * ks + t;
                hh_split(k0 ? ap[0] : 0.f, ah[i][0], al[i][0]);
                hh_split(k0 ? ap[8 * d] : 0.f, ah[i][1], al[i][1]);
                hh_split(k1 ? ap[4] : 0.f, ah[i][2], al[i][2]);
                hh_split(k1 ? ap[8 * d + 4] : 0.f, ah[i][3], al[i][3]);
            }
            const uint32_t* bh = Wh + (8 * ks + t) * NP + g;
            const uint32_t* bl = Wl + (8 * ks + t) * NP + g;
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                if (j < NT) {
                    const uint32_t bh0 = bh[8 * j], bh1 = bh[8 * j + 4 * NP], bl0 = bl[8 * j], bl1 = bl[8 * j + 4 * NP];
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        hh_mma(acc[i][j], al[i], bh0, bh1);
                        hh_mma(acc[i][j], ah[i], bl0, bl1);
                        hh_mma(acc[i][j], ah[i], bh0, bh1);
                    }
                }
            }
        }
        __syncwarp();                                  // every lane is done reading the range: stage the products in its place
#pragma unroll
        for (int j = 0; j < NTC; ++j) {
            const int col = 8 * j + 2 * t;
            if (j < NT && col < d) {
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    float* r0 = Xb + (16 * i + g) * d + col;
                    float* r1 = r0 + 8 * d;
                    r0[0] = acc[i][j][0]; r1[0] = acc[i][j][2];
                    if (col + 1 < d) { r0[1] = acc[i][j][1]; r1[1] = acc[i][j][3]; }
                }
            }
        }
        __syncwarp();
        float* gy = y + rng * RW * d;
        const int n4 = vec_out ? n >> 2 : 0;
        for (int i = lane; i < n4; i += 32) __stcs(reinterpret_cast<float4*>(gy) + i, *reinterpret_cast<const float4*>(Xb + 4 * i));
        for (int i = 4 * n4 + lane; i < n; i += 32) __stcs(gy + i, Xb[i]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
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
    const int rows_per_cta = 4 * kHhRowsPerWarp;       // 4 warps
    hh_matrix_kernel<<<(d + rows_per_cta - 1) / rows_per_cta, 128, 0, st>>>(Vs, n, d, W); HINT_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t hh_matrix_backward(const float* Vs, const float* W, const float* dW, int n, int d, float* dVs, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || n < 0) return cudaErrorInvalidValue;
    hh_matrix_bwd_kernel<<<1, 32 * kHhBwdWarps, 0, st>>>(Vs, W, dW, n, d, dVs); HINT_LAUNCHED();
    return cudaGetLastError();
}

namespace {
template <int MT, int NTC>
cudaError_t hh_apply_launch(const float* x, const float* W, long long B, int d, int transpose, float* y, cudaStream_t st) {
    const int K8 = (d + 7) & ~7, NP = K8 + ((K8 & 15) == 8 ? 0 : 8), RW = 16 * MT, bufsz = (RW * d + 8 + 3) & ~3;
    const size_t wbytes = sizeof(float) * 2 * (size_t)K8 * NP, per_warp = sizeof(float) * 2 * (size_t)bufsz;
    int nw = (int)(((size_t)216 * 1024 - wbytes) / per_warp);
    nw = nw > kHhApplyThreads / 32 ? kHhApplyThreads / 32 : nw < 1 ? 1 : nw;
    const size_t smem = wbytes + nw * per_warp;
    const void* fn = (const void*)hh_apply_kernel<MT, NTC>;
    // CTAs per SM for this d: asked once (the attribute call + occupancy query cost ~10 us of host time per launch otherwise)
    static std::atomic<int> cached[kHhMaxD + 1];
    int per_sm = cached[d].load(std::memory_order_relaxed);
    if (per_sm < 1) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * nw, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        cached[d].store(per_sm, std::memory_order_relaxed);
    }
    const long long nranges = (B + RW - 1) / RW, want = (nranges + nw - 1) / nw, cap = (long long)hh_sms() * per_sm;
    hh_apply_kernel<MT, NTC><<<(int)(want < cap ? want : cap), 32 * nw, smem, st>>>(x, W, B, d, transpose, y); HINT_LAUNCHED();
    return cudaGetLastError();
}
}  // namespace

cudaError_t hh_apply(const float* x, const float* W, long long B, int d, int transpose, float* y, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || B < 0) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    return d <= 48 ? hh_apply_launch<2, 6>(x, W, B, d, transpose, y, st) : hh_apply_launch<1, 16>(x, W, B, d, transpose, y, st);
}

size_t hh_wgrad_workspace_bytes(int d) { return mc_xt_y_workspace_bytes(d, d); }

// dW = x^T dz on the couplings' weight-gradient kernel (3 x TF32 MMAs, the samples split over warps and CTAs, fixed-order
// reductions): d = 43, 2^20 rows 1.16 ms (FFMA register-tile version) -> see profiles/ncu_r02_hh_apply.txt
cudaError_t hh_wgrad(const float* x, const float* dz, long long B, int d, float* dW, void* workspace, cudaStream_t st) {
    if (d < 1 || d > kHhMaxD || B < 0) return cudaErrorInvalidValue;
    return mc_xt_y(x, dz, d, d, B, dW, workspace, hh_wgrad_workspace_bytes(d), st);
}

}  // namespace hint
