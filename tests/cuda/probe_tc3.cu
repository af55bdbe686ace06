// Hardware probe for the primitives of the tcgen05 training kernel (tc3), run on the B200 via gpurun:
//   T1  weight-gradient style SS GEMM  D[feature j][feature i] = sum_s A[j][s] * B[i][s]  with both operands written by
//       "epilogue" threads (thread = sample) into K-major images (K = samples): SWIZZLE_128B (layout type 2) and the
//       un-swizzled canonical layout with a padded K stride (LBO = 144 B); second M tile (rows 128..255 of the image)
//   T2  K split of a skinny GEMM over 4 issuing warps with 4 separate accumulators (correctness + cycles)
//   T3  cycles of the GEMM shapes the planner has to budget: SS N=144/K=128, SS N=16/K=128 (1 and 4 issuers),
//       TS N=128/K=128, TS N=16/K=128 (1 and 4 issuers)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

__host__ __device__ inline int sw128_off(int r, int k, int rows) {   // float index; rows multiple of 8
    return (k >> 5) * rows * 32 + (r >> 3) * 256 + (r & 7) * 32 + ((((k & 31) >> 2) ^ (r & 7)) << 2) + (k & 3);
}
__host__ __device__ inline int pad_off(int r, int k) {               // type 0, LBO = 144 B, SBO = 32 * 144 B
    return (r >> 3) * (32 * 36) + (k >> 2) * 36 + (r & 7) * 4 + (k & 3);
}
__device__ inline uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ inline float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// layout 0: SW128 ; layout 1: padded type 0.  A has 192 rows (the second M tile reads 64 rows past it), B has NB rows; K = 128 samples.
__global__ void __launch_bounds__(256) t1_kernel(const float* gA, const float* gB, float* D, int NB, int layout, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* imgA = reinterpret_cast<float*>(smem);                 // 192 rows x 128 samples (+padding) = 110592 B
    float* imgB = reinterpret_cast<float*>(smem + 110592);        // up to 160 rows (+padding)
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    // "epilogue" writes: thread = sample (tid & 127); the two halves of the CTA split the rows
    {
        const int s = tid & 127, half = tid >> 7;
        for (int r = half; r < 192; r += 2) imgA[layout == 0 ? sw128_off(r, s, 192) : pad_off(r, s)] = gA[r * 128 + s];
        for (int r = half; r < NB; r += 2) imgB[layout == 0 ? sw128_off(r, s, 160) : pad_off(r, s)] = gB[r * 128 + s];
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, NB);
        long long t0 = clock64();
        for (int mt = 0; mt < 2; ++mt)
            for (int kk = 0; kk < 16; ++kk) {
                uint64_t da, db;
                if (layout == 0) {
                    da = desc_sw128(smem_u32(imgA) + (kk >> 2) * 192 * 128 + mt * 16 * 1024 + (kk & 3) * 32);
                    db = desc_sw128(smem_u32(imgB) + (kk >> 2) * 160 * 128 + (kk & 3) * 32);
                } else {
                    da = smem_desc(smem_u32(imgA) + mt * 16 * 4608 + kk * 288, 144, 4608);
                    db = smem_desc(smem_u32(imgB) + kk * 288, 144, 4608);
                }
                mma_ss(tb + mt * 160, da, db, idesc, kk > 0);
            }
        commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        cyc[0] = t1 - t0; cyc[1] = t2 - t0;
    } else {
        mbar_wait(&bar, 0);
    }
    fence_after_sync();
    if (warp < 4)
        for (int mt = 0; mt < 2; ++mt)
            for (int c = 0; c < NB; c += 8) {
                float v[8];
                ld8(taddr(tb, warp * 32, mt * 160 + c), v);
                wait_ld();
                for (int j = 0; j < 8; ++j) D[(mt * 128 + tid) * 160 + c + j] = v[j];
            }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// T2/T3: timing + K split.  mode: 0 SS, 1 TS.  W issuing warps (warps 4..4+W-1) split the 16 K steps; each has its own
// accumulator at column 256 + iw*N (W*N <= 256).  A (SS) = SW128 image rows 0..127; A (TS) = TMEM columns 0..127.
__global__ void __launch_bounds__(256) t3_kernel(const float* gA, const float* gB, float* D, int N, int mode, int W, int reps, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* imgA = reinterpret_cast<float*>(smem);                 // 128 rows x 128 samples SW128 = 64 KB
    float* imgB = reinterpret_cast<float*>(smem + 65536);         // SS: SW128 image N rows; TS: canonical K-major [N][128]
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    {
        const int s = tid & 127, half = tid >> 7;
        for (int r = half; r < 128; r += 2) imgA[sw128_off(r, s, 128)] = gA[r * 128 + s];
        for (int r = half; r < N; r += 2) {
            if (mode == 0) imgB[sw128_off(r, s, 256)] = gB[r * 128 + s];
            else imgB[canon_off(r, s, 128)] = gB[r * 128 + s];
        }
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    if (mode == 1 && warp < 4) {   // TS: A[m = sample lane][k = column]  (A^T of the SS case: D = A(lanes) * B^T)
        for (int c = 0; c < 128; c += 8) {
            float v[8];
            for (int j = 0; j < 8; ++j) v[j] = gA[tid * 128 + c + j];
            st8(taddr(tb, warp * 32, c), v);
        }
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int iw = warp - 4;
    if (iw >= 0 && iw < W && elect_one()) {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t dcol = tb + 256 + iw * N;
        const int k0 = iw * (16 / W), k1 = k0 + 16 / W;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kk = k0; kk < k1; ++kk) {
                if (mode == 0) {
                    const uint64_t da = desc_sw128(smem_u32(imgA) + (kk >> 2) * 128 * 128 + (kk & 3) * 32);
                    const uint64_t db = desc_sw128(smem_u32(imgB) + (kk >> 2) * 256 * 128 + (kk & 3) * 32);
                    mma_ss(dcol, da, db, idesc, (kk > k0) || r > 0);
                } else {
                    const uint64_t db = smem_desc(smem_u32(imgB) + kk * 256, 128, 128 * 32);
                    mma_ts(dcol, tb + kk * 8, db, idesc, (kk > k0) || r > 0);
                }
            }
        commit(&bars[iw]);
        mbar_wait(&bars[iw], 0);
        long long t2 = clock64();
        cyc[iw] = t2 - t0;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp < 4)
        for (int w2 = 0; w2 < W; ++w2)
            for (int c = 0; c < N; c += 8) {
                float v[8];
                ld8(taddr(tb, warp * 32, 256 + w2 * N + c), v);
                wait_ld();
                for (int j = 0; j < 8; ++j) D[(w2 * 128 + tid) * 256 + c + j] = v[j];
            }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

static float tf(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(256 * 128), B(256 * 128);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dB, *dD; long long* dc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 512 * 256 * 4); cudaMalloc(&dc, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    if (cudaFuncSetAttribute(t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) != cudaSuccess) printf("attr t1 failed\n");
    if (cudaFuncSetAttribute(t3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) != cudaSuccess) printf("attr t3 failed\n");
    std::vector<float> D(512 * 256);
    printf("== T1: SS weight-gradient GEMM, operands written by sample-threads into K-major images (K = 128 samples), 2 M tiles\n");
    for (int layout = 0; layout < 2; ++layout)
        for (int NB : {16, 144, 160}) {
            cudaMemset(dD, 0, 512 * 256 * 4);
            t1_kernel<<<1, 256, 110592 + 160 * 576 + 1024>>>(dA, dB, dD, NB, layout, dc);
            cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
            long long cyc[2] = {0, 0};
            cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, 256 * 160 * 4, cudaMemcpyDeviceToHost);
            double err = 0, ref_max = 0;
            for (int j = 0; j < 192; ++j)
                for (int i = 0; i < NB; ++i) {
                    double acc = 0;
                    for (int s = 0; s < 128; ++s) acc += (double)tf(A[j * 128 + s]) * tf(B[i * 128 + s]);
                    err = fmax(err, fabs(acc - D[j * 160 + i])); ref_max = fmax(ref_max, fabs(acc));
                }
            printf("layout %s NB=%3d : max abs err %.3e (ref max %.2f)  issue %lld cyc, complete %lld cyc for 32 MMAs [%s]\n",
                   layout == 0 ? "SW128 " : "pad144", NB, err, ref_max, cyc[0], cyc[1], cudaGetErrorString(e));
            if (e != cudaSuccess) return 0;
        }
    printf("== T2/T3: 16 K steps (K = 128) x reps, W issuing warps splitting K into separate accumulators\n");
    const int reps = 8;
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {16, 32, 64, 128})
            for (int W : {1, 2, 4}) {
                if (W * N > 256) continue;
                cudaMemset(dD, 0, 512 * 256 * 4);
                t3_kernel<<<1, 256, 65536 + 131072 + 1024>>>(dA, dB, dD, N, mode, W, reps, dc);
                cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
                long long cyc[4] = {0, 0, 0, 0};
                cudaMemcpy(cyc, dc, 32, cudaMemcpyDeviceToHost);
                cudaMemcpy(D.data(), dD, 512 * 256 * 4, cudaMemcpyDeviceToHost);
                double err = 0;
                for (int j = 0; j < 128; ++j)
                    for (int i = 0; i < N; ++i) {
                        double acc = 0, got = 0;
                        for (int s = 0; s < 128; ++s)
                            acc += mode == 0 ? (double)tf(A[j * 128 + s]) * tf(B[i * 128 + s]) : (double)tf(A[j * 128 + s]) * tf(B[i * 128 + s]);
                        for (int w2 = 0; w2 < W; ++w2) got += D[(w2 * 128 + j) * 256 + i];
                        err = fmax(err, fabs(acc * reps - got));
                    }
                long long mx = 0; for (int i = 0; i < W; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
                printf("%s N=%3d W=%d : max abs err %.3e   %lld cyc for %d MMAs -> %.1f cyc/MMA aggregate [%s]\n", mode == 0 ? "SS" : "TS", N, W, err,
                       mx, 16 * reps, (double)mx / (16 * reps), cudaGetErrorString(e));
                if (e != cudaSuccess) return 0;
            }
    return 0;
}
