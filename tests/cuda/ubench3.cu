// Micro-benchmarks that size the tcgen05 design (run on the B200 via gpurun):
//   S1  tcgen05.mma.kind::tf32 issue cost / throughput with WARP-UNIFORM operands (no waterfall loop), TS and SS,
//       N = 16..256, one accumulator chain or four, 1/2/4 issuing warps
//   S2  TMEM load / store throughput (ld16, relu, st16 round trips) with 4 / 8 / 16 warps
//   S3  correctness of MN-major tf32 operands (needed by the weight-gradient GEMMs, which contract over samples)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

// ------------------------------------------------------------------------------------------------ S1
// ts: A from TMEM; chains: number of distinct accumulators rotated through; W issuing warps (warps 4..4+W-1)
__global__ void __launch_bounds__(384) s1_kernel(long long* out, int N, int ts, int chains, int W, int nouter) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sB = reinterpret_cast<float*>(smem);
    for (int i = tid; i < 131072 / 4; i += 384) sB[i] = 0.001f * (i % 97);
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (slot != 0) { if (tid == 0) out[0] = -1; return; }   // base must be 0 for the uniform-operand variant
    if (warp < 4) {
        float v[16];
        for (int e = 0; e < 16; ++e) v[e] = 0.5f;
        for (int c = 0; c < 512; c += 16) st16(((uint32_t)(warp * 32) << 16) + c, v);
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int iw = warp - 4;
    if (iw >= 0 && iw < W) {
        // everything below depends only on kernel parameters and the warp index -> provably warp-uniform
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t sbase = smem_u32(smem);
        const uint32_t dbase = 64 + (uint32_t)iw * 112;           // per-warp accumulator region (N*chains <= 112 unless W == 1)
        const uint32_t bhi = (8 * 16) | (1u << 14);
        const uint32_t blo0 = (sbase >> 4) | (8u << 16);
        const uint32_t ahi = (8 * 16) | (1u << 14);
        const uint32_t alo0 = ((sbase + 71680) >> 4) | (8u << 16);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int o = 0; o < nouter; ++o) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t d = dbase + (uint32_t)((k % chains) * N);
                    const uint64_t bd = ((uint64_t)bhi << 32) | (blo0 + (uint32_t)(k & 7) * 16);
                    if (ts) mma_ts(d, (uint32_t)((k & 7) * 8), bd, idesc, 1u);
                    else mma_ss(d, ((uint64_t)ahi << 32) | (alo0 + (uint32_t)(k & 7) * 16), bd, idesc, 1u);
                }
            }
            t1 = clock64();
            commit(&bars[iw]);
            mbar_wait(&bars[iw], 0);
            t2 = clock64();
            out[1 + iw * 2] = t1 - t0;
            out[2 + iw * 2] = t2 - t0;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(0, 512);
}

// ------------------------------------------------------------------------------------------------ S2
// op: 0 = ld16 only, 1 = st16 only, 2 = ld16 + relu + st16, 3 = ld32 only (two x16), 4 = ld16+relu+st16 with one wait per 4 groups
__global__ void __launch_bounds__(512) s2_kernel(long long* out, float* sink, int nwarps, int op, int reps) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    if (warp < 4) {
        float v[16];
        for (int e = 0; e < 16; ++e) v[e] = 0.5f - (float)((tid + e) & 1);
        for (int c = 0; c < 512; c += 16) st16(tb + lane_base + c, v);
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int per = nwarps / 4;            // warps sharing one lane quadrant
    const int sub = warp >> 2;             // which column slice
    long long t0 = clock64();
    float acc = 0.f;
    if (warp < nwarps) {
        const int c0 = sub * (512 / per), c1 = c0 + 512 / per;
        for (int r = 0; r < reps; ++r) {
            if (op == 5) {
                for (int c = c0; c < c1; c += 64) {
                    float v0[16], v1[16], v2[16], v3[16];
                    ld16(tb + lane_base + c, v0); ld16(tb + lane_base + c + 16, v1);
                    ld16(tb + lane_base + c + 32, v2); ld16(tb + lane_base + c + 48, v3);
                    wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc += fmaxf(v0[e], v1[e]) + fmaxf(v2[e], v3[e]);
                }
                continue;
            }
            if (op == 6) {
                for (int c = c0; c < c1; c += 64) {
                    float v0[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v0[e] = (float)(r + e + c);
                    st16(tb + lane_base + c, v0); st16(tb + lane_base + c + 16, v0);
                    st16(tb + lane_base + c + 32, v0); st16(tb + lane_base + c + 48, v0);
                }
                wait_st();
                continue;
            }
            if (op == 4) {
                for (int c = c0; c < c1; c += 64) {
                    float v0[16], v1[16], v2[16], v3[16];
                    ld16(tb + lane_base + c, v0); ld16(tb + lane_base + c + 16, v1);
                    ld16(tb + lane_base + c + 32, v2); ld16(tb + lane_base + c + 48, v3);
                    wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) { v0[e] = fmaxf(v0[e], 0.f); v1[e] = fmaxf(v1[e], 0.f); v2[e] = fmaxf(v2[e], 0.f); v3[e] = fmaxf(v3[e], 0.f); }
                    st16(tb + lane_base + c, v0); st16(tb + lane_base + c + 16, v1);
                    st16(tb + lane_base + c + 32, v2); st16(tb + lane_base + c + 48, v3);
                }
                wait_st();
                continue;
            }
            for (int c = c0; c < c1; c += 16) {
                float v[16];
                if (op == 0 || op == 2 || op == 3) {
                    ld16(tb + lane_base + c, v);
                    wait_ld();
                    if (op != 2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc += v[e];
                    }
                }
                if (op == 1) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = (float)(r + e);
                }
                if (op == 2) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (op == 1 || op == 2) st16(tb + lane_base + c, v);
            }
            if (op == 1 || op == 2) wait_st();
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[tid] = acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------ S3
// D[128 x N] = A[128 x K] * B[N x K]^T with A and/or B stored MN-major (M / N index contiguous), no swizzle.
// element (m, k) at  (m & 3) + (k & 7) * 4 + (m >> 2) * 32 + (k >> 3) * (R / 4) * 32     [floats], R = rows (128 or N)
// variant 0: MN stride (128 B) in the SBO field, K stride in LBO; variant 1: swapped.
__host__ __device__ inline int mn_off(int m, int k, int R) { return (m & 3) + (k & 7) * 4 + (m >> 2) * 32 + (k >> 3) * (R / 4) * 32; }

__global__ void __launch_bounds__(128) s3_kernel(const float* __restrict__ A, const float* __restrict__ B, float* D, int K, int N,
                                                 int a_mn, int b_mn, int variant, unsigned* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + 128 * K;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 128 * K; i += 128) { int r = i / K, k = i % K; sA[a_mn ? mn_off(r, k, 128) : canon_off(r, k, K)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; sB[b_mn ? mn_off(r, k, N) : canon_off(r, k, K)] = B[i]; }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    const uint32_t idesc = idesc_tf32(128, N, a_mn, b_mn);
    if (tid == 0) {
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad, bd;
            if (a_mn) {
                const uint32_t kstride = (128 / 4) * 128, mnstride = 128;
                ad = variant == 0 ? smem_desc(smem_u32(sA) + ks * kstride, kstride, mnstride) : smem_desc(smem_u32(sA) + ks * kstride, mnstride, kstride);
            } else ad = smem_desc(smem_u32(sA) + ks * 256, 128, (K / 4) * 128);
            if (b_mn) {
                const uint32_t kstride = (N / 4) * 128, mnstride = 128;
                bd = variant == 0 ? smem_desc(smem_u32(sB) + ks * kstride, kstride, mnstride) : smem_desc(smem_u32(sB) + ks * kstride, mnstride, kstride);
            } else bd = smem_desc(smem_u32(sB) + ks * 256, 128, (K / 4) * 128);
            mma_ss(tb, ad, bd, idesc, ks > 0);
        }
        commit(&bar);
    }
    unsigned st = 0;
    if (!mbar_wait_bounded(&bar, 0, 1u << 22)) st = 1;
    fence_after_sync();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        ld8(taddr(tb, warp * 32, c), v);
        wait_ld();
        for (int j = 0; j < 8; ++j) D[tid * N + c + j] = v[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
    if (st) atomicOr(status, st);
}

static float tf32_trunc(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int main(int argc, char** argv) {
    const bool only_s2 = argc > 1;
    long long* d; cudaMalloc(&d, 256);
    float* sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    printf("== S1: tcgen05.mma tf32 M=128 K=8, uniform operands; cycles per MMA: issue-only / until complete (tensor floor N/2)\n");
    for (int ts = 1; ts >= 0 && !only_s2; --ts)
        for (int N : {16, 32, 64, 128, 256})
            for (int chains : {1, 4})
                for (int W : {1, 2, 4}) {
                    if (N * chains > 448 && W == 1) continue;
                    if (W > 1 && N * chains > 112) continue;
                    const int nouter = 8;
                    cudaMemset(d, 0, 256);
                    s1_kernel<<<1, 384, 131072>>>(d, N, ts, chains, W, nouter);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
                    long long iss = 0, tot = 0;
                    for (int w = 0; w < W; ++w) { if (h[1 + 2 * w] > iss) iss = h[1 + 2 * w]; if (h[2 + 2 * w] > tot) tot = h[2 + 2 * w]; }
                    const int nm = 16 * nouter;
                    printf("%s N=%3d chains=%d W=%d : issue %6.1f  complete %6.1f cyc per MMA per warp  (aggregate %5.1f cyc/MMA, floor %d) [%s]%s\n",
                           ts ? "TS" : "SS", N, chains, W, (double)iss / nm, (double)tot / nm, (double)tot / (nm * W), N / 2,
                           cudaGetErrorString(e), h[0] == -1 ? " TMEM base != 0" : "");
                }
    printf("== S2: TMEM ld/st throughput, 512 columns x 128 lanes per rep; cycles per rep and bytes/cycle/SM\n");
    for (int op : {4, 5, 6})
        for (int nw : {4, 8, 16}) {
            const int reps = 20;
            s2_kernel<<<1, 512>>>(d, sink, nw, op, reps);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[1]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            const double cyc = (double)h[0] / reps;
            const char* names[] = {"ld16", "st16", "ld16+relu+st16", "", "4x(ld16) wait 4x(st16)", "4x(ld16) wait", "4x(st16)"};
            printf("%-24s warps=%2d : %8.1f cyc per 512 cols  -> %6.1f elements/cyc  (%6.1f B/cyc each way) [%s]\n", names[op], nw, cyc,
                   512.0 * 128 / cyc, 512.0 * 128 * 4 / cyc, cudaGetErrorString(e));
        }
    printf("== S3: MN-major tf32 operands (no swizzle)\n");
    if (!only_s2) {
        const int K = 16, N = 32;
        std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
        srand(1);
        for (auto& v : A) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
        for (auto& v : B) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
        float *dA, *dB, *dD; unsigned* dst;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dst, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(s3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn) for (int variant = 0; variant < 2; ++variant) {
            if (!amn && !bmn && variant) continue;
            cudaMemset(dD, 0, D.size() * 4); cudaMemset(dst, 0, 4);
            s3_kernel<<<1, 128, 65536>>>(dA, dB, dD, K, N, amn, bmn, variant, dst);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned st; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double err = 0; for (size_t i = 0; i < D.size(); ++i) err = fmax(err, fabs((double)D[i] - R[i]));
            printf("A %s  B %s  variant %d : max abs err %.3e  status %u [%s]\n", amn ? "MN-major" : "K-major ", bmn ? "MN-major" : "K-major ", variant, err, st, cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("aborting S3 after CUDA error\n"); return 0; }
        }
    }
    return 0;
}
