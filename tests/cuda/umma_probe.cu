// Probe of the tcgen05 building blocks used by the TF32 kernels (run on the B200 via gpurun):
//   1. SS MMA: A, B from shared memory in the un-swizzled K-major canonical layout, D in TMEM
//   2. TS MMA: A from TMEM (written with tcgen05.st at an arbitrary column offset), B from shared memory
//   3. chain : relu(D) written back to TMEM and used as the A operand of a second MMA
// Every wait is bounded, so a wrong descriptor shows up as "timeout", not as a hung GPU.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"

using namespace hint::tc;

struct Params { int K, N, a_col, d_col, d2_col; };

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ B2,
                                             float* D_ss, float* D_ts, float* D_chain, Params p, unsigned* status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = p.K, N = p.N;
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + 128 * K;
    float* sB2 = sB + N * K;           // [N x N]: second layer weights (K2 = N)
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 128 * K; i += 128) { int r = i / K, k = i % K; sA[canon_off(r, k, K)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; sB[canon_off(r, k, K)] = B[i]; }
    for (int i = tid; i < N * N; i += 128) { int r = i / N, k = i % N; sB2[canon_off(r, k, N)] = B2[i]; }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = idesc_tf32(128, N);
    uint32_t phase = 0;
    unsigned st = 0;

    // ---- 1. SS ----
    if (tid == 0) {
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad = smem_desc(smem_u32(sA) + ks * 256, 128, (K / 4) * 128);
            uint64_t bd = smem_desc(smem_u32(sB) + ks * 256, 128, (K / 4) * 128);
            mma_ss(tbase + p.d_col, ad, bd, idesc, ks > 0);
        }
        commit(&bar);
    }
    if (!mbar_wait_bounded(&bar, phase, 1u << 22)) st |= 1;
    phase ^= 1;
    fence_after_sync();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        ld8(taddr(tbase, warp * 32, p.d_col + c), v);
        wait_ld();
        for (int j = 0; j < 8; ++j) D_ss[tid * N + c + j] = v[j];
    }
    fence_before_sync();
    __syncthreads();

    // ---- 2. TS: A rows -> TMEM columns [a_col, a_col+K) ----
    for (int c = 0; c < K; c += 8) {
        float v[8];
        for (int j = 0; j < 8; ++j) v[j] = A[tid * K + c + j];
        st8(taddr(tbase, warp * 32, p.a_col + c), v);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t bd = smem_desc(smem_u32(sB) + ks * 256, 128, (K / 4) * 128);
            mma_ts(tbase + p.d2_col, tbase + p.a_col + ks * 8, bd, idesc, ks > 0);
        }
        commit(&bar);
    }
    if (!mbar_wait_bounded(&bar, phase, 1u << 22)) st |= 2;
    phase ^= 1;
    fence_after_sync();
    // read D, store, and write relu(D) back IN PLACE as the A operand of the chained MMA
    for (int c = 0; c < N; c += 8) {
        float v[8];
        ld8(taddr(tbase, warp * 32, p.d2_col + c), v);
        wait_ld();
        for (int j = 0; j < 8; ++j) { D_ts[tid * N + c + j] = v[j]; v[j] = fmaxf(v[j], 0.f); }
        st8(taddr(tbase, warp * 32, p.d2_col + c), v);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    // ---- 3. chain: D3 = relu(D2) * B2^T, K2 = N ----
    if (tid == 0) {
        for (int ks = 0; ks < N / 8; ++ks) {
            uint64_t bd = smem_desc(smem_u32(sB2) + ks * 256, 128, (N / 4) * 128);
            mma_ts(tbase + p.d_col, tbase + p.d2_col + ks * 8, bd, idesc, ks > 0);
        }
        commit(&bar);
    }
    if (!mbar_wait_bounded(&bar, phase, 1u << 22)) st |= 4;
    phase ^= 1;
    fence_after_sync();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        ld8(taddr(tbase, warp * 32, p.d_col + c), v);
        wait_ld();
        for (int j = 0; j < 8; ++j) D_chain[tid * N + c + j] = v[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
    if (st) atomicOr(status, st);
}

static float tf32r(float x) {  // round-to-nearest-even-ish emulation of dropping 13 mantissa bits (truncate variant checked too)
    unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y;
}

int main() {
    int dev_ok = 0; cudaGetDeviceCount(&dev_ok);
    if (!dev_ok) { printf("no device\n"); return 1; }
    const int cases[][5] = {  // K, N, a_col, d_col, d2_col
        {8, 16, 64, 0, 128}, {24, 32, 64, 0, 128}, {24, 80, 68, 0, 256}, {16, 144, 100, 160, 320}, {72, 80, 4, 96, 200},
        {24, 48, 65, 0, 128}, {8, 16, 66, 16, 48}, {40, 24, 64, 0, 128},
    };
    for (auto& cs : cases) {
        Params p{cs[0], cs[1], cs[2], cs[3], cs[4]};
        const int K = p.K, N = p.N;
        std::vector<float> A(128 * K), B(N * K), B2(N * N);
        srand(1234 + K * 7 + N);
        for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
        for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
        for (auto& v : B2) v = (rand() % 2001 - 1000) / 1000.f;
        float *dA, *dB, *dB2, *dss, *dts, *dch; unsigned* dst;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dB2, B2.size() * 4);
        cudaMalloc(&dss, 128 * N * 4); cudaMalloc(&dts, 128 * N * 4); cudaMalloc(&dch, 128 * N * 4); cudaMalloc(&dst, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dss, 0xFF, 128 * N * 4); cudaMemset(dts, 0xFF, 128 * N * 4); cudaMemset(dch, 0xFF, 128 * N * 4); cudaMemset(dst, 0, 4);
        size_t smem = (size_t)(128 * K + N * K + N * N) * 4 + 128;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<1, 128, smem>>>(dA, dB, dB2, dss, dts, dch, p, dst);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> ss(128 * N), ts(128 * N), ch(128 * N); unsigned st = 0;
        cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(ch.data(), dch, ch.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
        double e_ss = 0, e_ts = 0, e_ch = 0, e_full = 0;
        for (int r = 0; r < 128; ++r) {
            std::vector<double> h(N);
            for (int n = 0; n < N; ++n) {
                double acc = 0, accf = 0;
                for (int k = 0; k < K; ++k) { acc += (double)tf32r(A[r * K + k]) * tf32r(B[n * K + k]); accf += (double)A[r * K + k] * B[n * K + k]; }
                e_ss = fmax(e_ss, fabs(acc - ss[r * N + n]));
                e_ts = fmax(e_ts, fabs(acc - ts[r * N + n]));
                e_full = fmax(e_full, fabs(accf - ss[r * N + n]));
                h[n] = ts[r * N + n] > 0 ? ts[r * N + n] : 0;
            }
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < N; ++k) acc += (double)tf32r((float)h[k]) * tf32r(B2[n * N + k]);
                e_ch = fmax(e_ch, fabs(acc - ch[r * N + n]));
            }
        }
        printf("K=%3d N=%3d a_col=%3d d_col=%3d d2_col=%3d : cuda=%s status=%u  err SS %.3e (vs fp32 %.3e)  TS %.3e  chain %.3e\n",
               K, N, p.a_col, p.d_col, p.d2_col, cudaGetErrorString(e), st, e_ss, e_full, e_ts, e_ch);
        if (e != cudaSuccess) return 2;
        cudaFree(dA); cudaFree(dB); cudaFree(dB2); cudaFree(dss); cudaFree(dts); cudaFree(dch); cudaFree(dst);
    }
    return 0;
}
