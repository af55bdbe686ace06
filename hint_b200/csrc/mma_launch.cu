// Launch glue of the warp-MMA kernels: device tables, occupancy, kernel dispatch over the tile size / precision mode.
#include "mma_launch.h"

#include <algorithm>
#include <vector>

#include "mma_kernels.cuh"

namespace hint {

namespace {

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& v) {
    *dst = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

template <int TM, bool X3>
cudaError_t setup_kernels(const MSchedule& s, bool bwd, int num_sms, int* max_ctas) {
    const void* fn = bwd ? (const void*)hint_bwd_mma_kernel<TM, X3> : (const void*)hint_fwd_mma_kernel<TM, X3>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kMmaThreads, s.smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    *max_ctas = std::min(*max_ctas > 0 ? *max_ctas : occ * num_sms, occ * num_sms);
    return cudaSuccess;
}

cudaError_t setup_schedule(const MSchedule& s, bool bwd, int num_sms, DevMmaSchedule& ds) {
    cudaError_t e;
    if ((e = upload(&ds.prog, s.prog)) != cudaSuccess) return e;
    if ((e = upload(&ds.eps, s.eps)) != cudaSuccess) return e;
    ds.max_ctas = 0;
#define HINT_SETUP(TMV)                                                                                      \
    case TMV:                                                                                                \
        if ((e = setup_kernels<TMV, false>(s, bwd, num_sms, &ds.max_ctas)) != cudaSuccess) return e;         \
        return setup_kernels<TMV, true>(s, bwd, num_sms, &ds.max_ctas);
    switch (s.TM) {
        HINT_SETUP(64) HINT_SETUP(32) HINT_SETUP(16)
    }
#undef HINT_SETUP
    return cudaErrorInvalidValue;
}

MmaTables make_tables(const Plan& p, const MSchedule& s, const DevMmaSchedule& ds, int prog) {
    MmaTables t;
    t.prog = ds.prog; t.eps = ds.eps;
    for (int w = 0; w < kMmaWarps; ++w) t.begin[w] = s.prog_begin[prog][w];
    t.d = p.d; t.dc = p.dc;
    t.col_x = s.col_x; t.col_d = s.col_d; t.col_one = s.col_one; t.col_zero = s.col_zero;
    t.raw_off = s.raw_off;
    t.alpha = p.alpha;
    return t;
}

}  // namespace

cudaError_t mma_setup(const MmaPlan& m, int num_sms, DevMma& d) {
    cudaError_t e;
    if ((e = setup_schedule(m.fwd, false, num_sms, d.fwd)) != cudaSuccess) return e;
    if ((e = setup_schedule(m.bwd, true, num_sms, d.bwd)) != cudaSuccess) return e;
    if ((e = upload(&d.pack_src, m.pack_src)) != cudaSuccess) return e;
    return upload(&d.unpack_src, m.unpack_src);
}

void mma_free(DevMma& d) {
    for (DevMmaSchedule* s : {&d.fwd, &d.bwd}) { cudaFree(s->prog); cudaFree(s->eps); }
    cudaFree(d.pack_src); cudaFree(d.unpack_src);
}

cudaError_t mma_pack(const MmaPlan& m, const DevMma& d, const float* params, float* hi, float* lo, cudaStream_t st) {
    const long long n = m.n_packed;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
    hint_pack_mma_kernel<<<blocks, threads, 0, st>>>(d.pack_src, params, hi, lo, n);
    return cudaGetLastError();
}

cudaError_t mma_launch_fwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, const float* x, const float* c,
                           const float* hi, const float* lo, float* z, float* logdet, long long B, int rev, cudaStream_t st) {
    const MSchedule& s = m.fwd;
    const MmaTables T = make_tables(p, s, d.fwd, rev ? PROG_INV : PROG_FWD);
    const long long ntiles = (B + s.TM - 1) / s.TM;
    const int grid = (int)std::min<long long>(ntiles, d.fwd.max_ctas);
#define HINT_LAUNCH(TMV)                                                                                                   \
    case TMV:                                                                                                              \
        if (x3) hint_fwd_mma_kernel<TMV, true><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, x, c, hi, lo, z, logdet, B, rev ? 1 : 0);  \
        else hint_fwd_mma_kernel<TMV, false><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, x, c, hi, lo, z, logdet, B, rev ? 1 : 0);    \
        break;
    switch (s.TM) {
        HINT_LAUNCH(64) HINT_LAUNCH(32) HINT_LAUNCH(16)
        default: return cudaErrorInvalidValue;
    }
#undef HINT_LAUNCH
    return cudaGetLastError();
}

cudaError_t mma_launch_bwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, int grid, const float* z, const float* c,
                           const float* hi, const float* lo, const float* dz, const float* dlogdet, float* x_rec, float* dx,
                           float* dc, float* partials, long long B, cudaStream_t st) {
    const MSchedule& s = m.bwd;
    const MmaTables T = make_tables(p, s, d.bwd, PROG_BWD);
    const long long np = m.n_partial;
#define HINT_LAUNCH(TMV)                                                                                                   \
    case TMV:                                                                                                              \
        if (x3) hint_bwd_mma_kernel<TMV, true><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, z, c, hi, lo, dz, dlogdet, x_rec, dx, dc, partials, np, B);  \
        else hint_bwd_mma_kernel<TMV, false><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, z, c, hi, lo, dz, dlogdet, x_rec, dx, dc, partials, np, B);    \
        break;
    switch (s.TM) {
        HINT_LAUNCH(64) HINT_LAUNCH(32) HINT_LAUNCH(16)
        default: return cudaErrorInvalidValue;
    }
#undef HINT_LAUNCH
    return cudaGetLastError();
}

}  // namespace hint
