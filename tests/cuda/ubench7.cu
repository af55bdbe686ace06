// What does a tcgen05.commit between short runs of tcgen05.mma cost?  TS MMAs (A from TMEM, B canonical K-major slab in shared memory),
// N = 128, records of R MMAs, variants: no commit / commit after every record / commit + a try_wait on an already completed barrier.
#include <cuda_runtime.h>
#include <cstdio>
#include "../../hint_b200/csrc/tcgen05.cuh"
using namespace hint::tc;

__global__ void __launch_bounds__(576) bench(long long* out, int spin) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    for (int i = tid; i < 64 * 1024 / 4; i += 576) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    fence_proxy_async_smem();
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = idesc_tf32(128, 128);
        const uint32_t sb = smem_u32(smem);
        int k = 0;
        for (int variant = 0; variant < 4; ++variant) {
            for (int R : {1, 4, 8, 16}) {
                const int nrec = 64 / R * 4;     // 256 MMAs in total
                uint32_t ph[4] = {0, 0, 0, 0};
                const long long t0 = clock64();
                for (int r = 0; r < nrec; ++r) {
                    const uint32_t base = sb + (uint32_t)(r & 3) * 16384u;
                    if (variant == 3) { mbar_try_wait(&bars[7], 1); }   // a wait that succeeds immediately (phase 0 never started: parity 1 reads as complete)
                    for (int i = 0; i < R; ++i) {
                        const uint64_t db = smem_desc(base + (uint32_t)(i & 3) * 256u, 128, 1024);
                        mma_ts(256, (uint32_t)(i & 15) * 8, db, idesc, (r | i) ? 1u : 0u);
                    }
                    if (variant >= 1) commit(&bars[r & 3]);
                    if (variant == 2) { mbar_wait(&bars[r & 3], ph[r & 3]); ph[r & 3] ^= 1; }   // fully serialised: wait for completion
                }
                const long long t1 = clock64();
                commit(&bars[4]);
                mbar_wait(&bars[4], (uint32_t)(k & 1));
                const long long t2 = clock64();
                if (variant == 1 || variant == 3) {   // drain the per-record barriers' phases so later variants start clean
                }
                if (blockIdx.x == 0) { out[3 * k] = t1 - t0; out[3 * k + 1] = t2 - t0; out[3 * k + 2] = R * 100 + variant; }
                ++k;
                // re-arm: barriers 0..3 may have completed phases; reset by re-init (single thread, nothing in flight)
                for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
                fence_mbar_init();
            }
        }
        if (blockIdx.x == 0) out[63] = k;
        if (spin) { mbar_arrive(&bars[6]); }
    } else if (spin && warp >= 2) {
        // 16 warps polling an mbarrier for the whole run, like the epilogue warps of the training kernel while they wait for the tensor pipe
        mbar_wait(&bars[6], 0);
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int grid = (cfg & 1) ? 148 : 1, spin = cfg >> 1;
    for (int rep = 0; rep < 2; ++rep) {
        bench<<<grid, 576, 64 * 1024>>>(out, spin);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    }
    printf("---- grid %d, %s\n", grid, spin ? "16 warps polling an mbarrier" : "other warps idle");
    const char* names[] = {"no commit", "commit per record", "commit + wait completion per record", "commit per record + a passing try_wait"};
    for (int i = 0; i < out[63]; ++i)
        if (out[3 * i + 2] / 100 == 4) printf("R=%2lld %-42s issue %6.1f cyc/MMA   complete %6.1f cyc/MMA\n", out[3 * i + 2] / 100, names[out[3 * i + 2] % 100], out[3 * i] / 256.0, out[3 * i + 1] / 256.0);
  }
    return 0;
}
