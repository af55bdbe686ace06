// Legacy warp-level tensor-core path on B200: mma.sync.m16n8k8 tf32 throughput per SM (run via gpurun).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(1024) k(float* out, int iters, long long* cyc) {
    float c[4][4] = {};
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += c[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 4096;
        k<<<148, threads>>>(o, iters, c); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, threads>>>(o, iters, c); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        double flop_sm = 2.0 * 16 * 8 * 8 * 4 * iters * (threads / 32);
        printf("threads/SM %4d : %8.1f flop/cycle/SM   (%.1f TFLOP/s chip, %.3f ms) [%s]\n", threads, flop_sm / h, flop_sm * 148 / ms / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
