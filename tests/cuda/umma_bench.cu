// Micro-benchmark of tcgen05.mma (kind::tf32, A from TMEM, B from smem) issue/latency behaviour on B200:
// cycles for chains of small MMAs (dependent accumulate chain vs independent accumulators), and for the
// commit -> mbarrier -> wake round trip.  Results feed the scheduling model in DESIGN.md.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

__global__ void __launch_bounds__(160) bench(long long* out, int N, int nk, int nops, int independent, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* sB = reinterpret_cast<float*>(smem);
    for (int i = tid; i < 256 * 64; i += 160) sB[i] = 0.001f * (i % 97);
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    // zero TMEM so operands are finite
    if (warp < 4) {
        float v[16];
        for (int e = 0; e < 16; ++e) v[e] = 0.5f;
        for (int c = 0; c < 512; c += 16) st16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 128) {
        const uint32_t idesc = idesc_tf32(128, N);
        uint32_t phase = 0;
        long long best = 1ll << 60, best_issue = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64();
            for (int o = 0; o < nops; ++o) {
                const uint32_t d = tb + 256 + (independent ? (o * N) % 256 : 0);
                uint32_t blo = (smem_u32(sB) >> 4) | (8u << 16);
                const uint32_t bhi = (nk * 16) | (1u << 14);
                for (int ks = 0; ks < nk; ++ks) {
                    mma_ts(d, tb + ks * 8, ((uint64_t)bhi << 32) | blo, idesc, ks > 0);
                    blo += 16;
                }
            }
            long long t1 = clock64();
            commit(&bar);
            mbar_wait(&bar, phase);
            phase ^= 1;
            long long t2 = clock64();
            if (t2 - t0 < best) best = t2 - t0;
            if (t1 - t0 < best_issue) best_issue = t1 - t0;
        }
        out[0] = best; out[1] = best_issue;
        // empty commit round trip
        long long t0 = clock64();
        commit(&bar);
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        out[2] = t1 - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 64 * 4 + 1024);
    struct C { int N, nk, nops, ind; };
    std::vector<C> cases = {{16,1,1,0},{16,1,8,1},{16,1,32,1},{16,2,16,1},{16,9,1,0},{16,9,8,1},{32,9,1,0},{80,9,1,0},{80,9,2,1},{144,3,1,0},
                            {64,8,1,0},{64,8,4,1},{128,8,1,0},{256,8,1,0},{256,32,1,0},{16,1,64,0},{32,4,16,1}};
    for (auto c : cases) {
        bench<<<1, 160, 256 * 64 * 4 + 1024>>>(d, c.N, c.nk, c.nops, c.ind, 20);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        int nm = c.nk * c.nops;
        printf("N=%3d nk=%2d nops=%2d %s : total %6lld cyc (issue %5lld)  per-MMA %6.1f  ideal(N/2 per MMA) %5d  empty commit rt %lld  [%s]\n", c.N, c.nk, c.nops,
               c.ind ? "indep" : "chain", h[0], h[1], (double)h[0] / nm, nm * c.N / 2, h[2], cudaGetErrorString(e));
    }
    return 0;
}
