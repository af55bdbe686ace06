// Does tcgen05.mma issue scale with the number of issuing warps?  W warps, each one elected thread issuing its own
// chain of `nper` MMAs (N columns, independent accumulators), all committing to one mbarrier (count W).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

__global__ void __launch_bounds__(384) bench(long long* out, int N, int nper, int W, int lanes, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, go;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sB = reinterpret_cast<float*>(smem);
    for (int i = tid; i < 256 * 64; i += 384) sB[i] = 0.001f * (i % 97);
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, W * lanes); mbar_init(&go, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    if (warp < 4) {
        float v[16];
        for (int e = 0; e < 16; ++e) v[e] = 0.5f;
        for (int c = 0; c < 512; c += 16) st16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int iw = warp - 4;   // issuing warps 4..4+W-1
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        long long t0 = clock64();
        if (iw >= 0 && iw < W && lane < lanes) {
            const uint32_t idesc = idesc_tf32(128, N);
            const uint32_t d = tb + 128 + ((iw * lanes + lane) * N) % 384;
            uint32_t blo = (smem_u32(sB) >> 4) | (8u << 16);
            const uint32_t bhi = (8 * 16) | (1u << 14);
            for (int k = 0; k < nper; ++k) {
                mma_ts(d, tb + (k & 7) * 8, ((uint64_t)bhi << 32) | blo, idesc, k > 0);
                blo = (blo & 0xFFFF0000u) | ((blo + 16) & 0xFFFu);
            }
            commit(&bar);
        }
        if (tid == 128) {
            mbar_wait(&bar, r & 1);
            long long t2 = clock64();
            if (t2 - t0 < best) best = t2 - t0;
        }
    }
    if (tid == 128) out[0] = best;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 64 * 4 + 1024);
    struct C { int N, nper, W, lanes; };
    std::vector<C> cases = {{16,32,1,1},{16,16,2,1},{16,8,4,1},{16,4,8,1},{16,8,1,4},{16,4,1,8},{16,2,1,16},{16,1,1,32},{16,2,4,4},
                            {64,32,1,1},{64,8,4,1},{64,8,1,4},{128,32,1,1},{128,8,4,1},{128,8,1,4}};
    for (auto c : cases) {
        bench<<<1, 384, 256 * 64 * 4 + 1024>>>(d, c.N, c.nper, c.W, c.lanes, 20);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[1]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        int nm = c.nper * c.W * c.lanes;
        printf("N=%3d  %2d MMAs x %d warps x %2d lanes = %3d MMAs : total %6lld cyc  per-MMA %6.1f  tensor ideal %5d [%s]\n", c.N, c.nper, c.W, c.lanes, nm, h[0],
               (double)h[0] / nm, nm * c.N / 2, cudaGetErrorString(e));
    }
    return 0;
}
