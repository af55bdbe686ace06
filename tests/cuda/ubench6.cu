// Cycle cost of the synchronisation / TMEM primitives the tc3 epilogue uses per step (8 warps active, like the kernel).
#include <cuda_runtime.h>
#include <cstdio>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

#define T(name, code) { __syncthreads(); long long a = clock64(); _Pragma("unroll 1") for (int r = 0; r < 8; ++r) { code; } long long b = clock64(); if (tid == 0) out[k] = (b - a) / 8; ++k; }

__global__ void __launch_bounds__(256) bench(long long* out, const char** names, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bars[0], 8); mbar_init(&bars[1], 1); fence_mbar_init(); }
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t lb = (uint32_t)((warp & 3) * 32) << 16;
    float v[16]; for (int j = 0; j < 16; ++j) v[j] = tid + j;
    float* sm = reinterpret_cast<float*>(smem);
    int k = 0;
    uint32_t ph = 0;
    T("empty loop", asm volatile("" ::: "memory"));
    T("tcgen05.fence::after_thread_sync", fence_after_sync());
    T("tcgen05.fence::before_thread_sync", fence_before_sync());
    T("fence.proxy.async (nothing pending)", fence_proxy_async_smem());
    T("tcgen05.wait::st (nothing pending)", wait_st());
    T("tcgen05.wait::ld (nothing pending)", wait_ld());
    T("ld16 + wait_ld", ld16(lb + (warp >> 2) * 64, v); wait_ld());
    T("4 x ld16 + wait_ld", ld16(lb + (warp >> 2) * 64, v); ld16(lb + (warp >> 2) * 64 + 16, v); ld16(lb + (warp >> 2) * 64 + 32, v); ld16(lb + (warp >> 2) * 64 + 48, v); wait_ld());
    T("st16 + wait_st", st16(lb + (warp >> 2) * 64, v); wait_st());
    T("4 x st16 + wait_st", st16(lb + (warp >> 2) * 64, v); st16(lb + (warp >> 2) * 64 + 16, v); st16(lb + (warp >> 2) * 64 + 32, v); st16(lb + (warp >> 2) * 64 + 48, v); wait_st());
    T("st8 + wait_st", { float w[8]; for (int j = 0; j < 8; ++j) w[j] = v[j]; st8(lb + 256, w); wait_st(); });
    T("16 x STS (conflict free) ", for (int j = 0; j < 16; ++j) sm[j * 256 + tid] = v[j]);
    T("16 x STS + fence.proxy.async", for (int j = 0; j < 16; ++j) sm[j * 256 + tid] = v[j]; fence_proxy_async_smem());
    T("64 x STS + fence.proxy.async", for (int j = 0; j < 64; ++j) sm[j * 256 + tid] = v[j & 15]; fence_proxy_async_smem());
    T("__syncwarp + lane0 mbarrier.arrive (count 8) + all try_wait", { __syncwarp(); if (lane == 0) mbar_arrive(&bars[0]); mbar_wait(&bars[0], ph); ph ^= 1; });
    T("bar.sync 1, 256", named_bar_sync(1, 256));
    T("bar.sync (128 threads of a warpgroup)", named_bar_sync(2 + (warp >> 2), 128));
    T("full publish: wait_st, fence.proxy, fence::before, syncwarp, arrive", { wait_st(); fence_proxy_async_smem(); fence_before_sync(); __syncwarp(); if (lane == 0) mbar_arrive(&bars[0]); });
    T("5 x shfl.xor reduce", { float s = v[0]; for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh); v[0] = s; });
    T("ld.global.cg x16 independent (L2 hit) + add + st", { float o[16]; for (int j = 0; j < 16; ++j) o[j] = __ldcg(sink + 4096 + j * 128 + tid); for (int j = 0; j < 16; ++j) __stcg(sink + 4096 + j * 128 + tid, o[j] + 1.f); });
    if (k > 0 && tid == 0) out[63] = k;
    float acc = 0; for (int j = 0; j < 16; ++j) acc += v[j];
    sink[tid] = acc;
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

static const char* kNames[] = {"empty loop", "tcgen05.fence::after_thread_sync", "tcgen05.fence::before_thread_sync", "fence.proxy.async (nothing pending)", "tcgen05.wait::st (nothing pending)", "tcgen05.wait::ld (nothing pending)", "ld16 + wait_ld", "4 x ld16 + wait_ld", "st16 + wait_st", "4 x st16 + wait_st", "st8 + wait_st", "16 x STS (conflict free) ", "16 x STS + fence.proxy.async", "64 x STS + fence.proxy.async", "__syncwarp + lane0 mbarrier.arrive (count 8) + all try_wait", "bar.sync 1, 256", "bar.sync (128 threads of a warpgroup)", "full publish: wait_st, fence.proxy, fence::before, syncwarp, arrive", "5 x shfl.xor reduce", "ld.global.cg x16 independent (L2 hit) + add + st"};
int main() {
    long long* out; const char** names; float* sink;
    cudaMallocManaged(&out, 64 * 8); cudaMallocManaged(&names, 64 * 8); cudaMalloc(&sink, 1 << 20);
    cudaMemset(sink, 0, 1 << 20);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) { bench<<<1, 256, 100 * 1024>>>(out, names, sink); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } }
    for (int i = 0; i < out[63]; ++i) printf("%-70s %6lld cycles\n", kNames[i], out[i]);
    return 0;
}
