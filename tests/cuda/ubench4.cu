// Layout discovery for MN-major tf32 shared-memory operands of tcgen05.mma (run on the B200 via gpurun).
// The operand under test is an "address image": shared-memory float i holds the value i (exact in tf32 for i < 2048).
// The other operand is a one-hot selector, so D reports WHICH shared-memory float the tensor core used for each
// (row, k) of the operand.  Printed for several LBO / SBO values so the roles of the two strides can be read off.
//   mode 0: A under test (MN-major bit set), B = K-major one-hot  (D[m][n] = A(m, k=n), n < 8)
//   mode 1: B under test (MN-major bit set), A = K-major one-hot  (D[m][n] = B(n, k=m), m < 8)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

__global__ void __launch_bounds__(128) probe(float* D, int mode, int N, uint32_t lbo, uint32_t sbo, int tr_bit, int ltype, int hi, unsigned* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* img = reinterpret_cast<float*>(smem);            // 2048 floats: address image (8 KB)
    float* sel = img + 32768;                               // one-hot selector, K-major canonical, K = 8
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 32768; i += 128) img[i] = hi ? (float)(i >> 5) : (float)(i & 31);
    for (int i = tid; i < 256 * 8; i += 128) { int r = i / 8, k = i % 8; sel[canon_off(r, k, 8)] = (r == k) ? 1.f : 0.f; }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = slot;
    if (tid == 0) {
        const uint64_t dimg = smem_desc(smem_u32(img), lbo, sbo) | ((uint64_t)ltype << 61);
        const uint64_t dsel = smem_desc(smem_u32(sel), 128, 256);
        if (mode == 0) mma_ss(tb, dimg, dsel, idesc_tf32(128, N, tr_bit, 0), 0);
        else mma_ss(tb, dsel, dimg, idesc_tf32(128, N, 0, tr_bit), 0);
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

int main(int argc, char** argv) {
    const int tr = argc > 1 ? atoi(argv[1]) : 1;
    float* dD; unsigned* dst;
    cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dst, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    std::vector<float> D(128 * 256);
    struct Cfg { uint32_t lbo, sbo; };
    const Cfg cfgs[] = {{128, 256}, {256, 128}, {512, 128}, {128, 512}, {1024, 2048}, {2048, 1024}};
    const int ltypes[] = {1, 2, 4, 6, 0};
    for (int mode = 0; mode < 2; ++mode)
        for (int ltype : ltypes)
            for (const Cfg& c : cfgs) {
                const int N = 32;
                std::vector<float> lo(128 * N), hi(128 * N);
                unsigned st = 0; cudaError_t e = cudaSuccess;
                for (int h = 0; h < 2; ++h) {
                    cudaMemset(dD, 0, 128 * 256 * 4); cudaMemset(dst, 0, 4);
                    probe<<<1, 128, 160 * 1024>>>(dD, mode, N, c.lbo, c.sbo, tr, ltype, h, dst);
                    e = cudaDeviceSynchronize();
                    cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
                    cudaMemcpy(h ? hi.data() : lo.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) break;
                }
                printf("mode %d (%s under test) transpose=%d layout_type=%d LBO=%u SBO=%u status=%u [%s]\n", mode, mode == 0 ? "A" : "B", tr, ltype, c.lbo, c.sbo, st, cudaGetErrorString(e));
                if (e != cudaSuccess) return 0;
                const int R = mode == 0 ? 128 : N;
                for (int k = 0; k < 8; ++k) {
                    printf("  k=%d rows 0..%d:", k, R > 40 ? 39 : R - 1);
                    for (int r = 0; r < R && r < 40; ++r) { int idx = mode == 0 ? r * N + k : k * N + r; printf(" %5d", (int)hi[idx] * 32 + (int)lo[idx]); }
                    if (mode == 0) printf("  | r=64:%5d r=127:%5d", (int)hi[64 * N + k] * 32 + (int)lo[64 * N + k], (int)hi[127 * N + k] * 32 + (int)lo[127 * N + k]);
                    printf("\n");
                }
            }
    return 0;
}
