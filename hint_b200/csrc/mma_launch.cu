// Launch glue of the warp-MMA kernels: device tables, occupancy, kernel dispatch over the tile size / precision mode.
#include "mma_launch.h"
#include "launch_count.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
        HINT_SETUP(64) HINT_SETUP(32)
    }
#undef HINT_SETUP
    return cudaErrorInvalidValue;
}

MmaTables make_tables(const Plan& p, const MSchedule& s, const DevMmaSchedule& ds, int prog) {
    MmaTables t;
    t.prog = ds.prog; t.eps = ds.eps;
    t.in_param = s.fits_param[prog] ? 1 : 0;
    for (int w = 0; w < kMmaWarps; ++w) t.begin[w] = s.prog_begin[prog][w] - (t.in_param ? s.prog_begin[prog][0] : 0);
    t.d = p.d; t.dc = p.dc;
    t.col_x = s.col_x; t.col_d = s.col_d; t.col_one = s.col_one; t.col_zero = s.col_zero;
    t.raw_off = s.raw_off;
    t.alpha = p.alpha;
    t.dbg = nullptr;
    t.wcopies = 1; t.wstride = 0;
    return t;
}

// HINT_B200_MMA_DEBUG=1: per-barrier clock64 stamps of CTA 0's second tile, printed after a sync (developer aid)
long long* g_dbg = nullptr;
const char* dev_getenv(const char* name) {   // developer switches exist only in -DHINT_B200_DEV builds
#ifdef HINT_B200_DEV
    return std::getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}
bool dbg_on() { static const bool on = dev_getenv("HINT_B200_MMA_DEBUG") != nullptr; return on; }
void dbg_begin(MmaTables& T, cudaStream_t st) {
    if (!dbg_on()) return;
    if (!g_dbg) cudaMalloc((void**)&g_dbg, 1001 * sizeof(long long));
    cudaMemsetAsync(g_dbg, 0, 1001 * sizeof(long long), st);
    T.dbg = g_dbg;
}
void dbg_end(const char* what, cudaStream_t st) {
    if (!dbg_on()) return;
    static long long h[1001];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, g_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[hint_b200 mma dbg] %s: %lld stamps; cycles between barriers:", what, h[0]);
    for (long long i = 1; i < h[0]; ++i) std::fprintf(stderr, " %lld", h[1 + i] - h[i]);
    std::fprintf(stderr, "\n");
}

void fill_param_prog(const MSchedule& s, int prog, MmaParamProg& P) {
    std::memset(&P, 0, sizeof(P));
    if (!s.fits_param[prog]) return;
    const int b = s.prog_begin[prog][0], n = s.prog_end[prog] - b;
    std::memcpy(P.ops, s.prog.data() + b, (size_t)n * sizeof(WOp));
    for (size_t i = 0; i < s.eps.size(); ++i) {
        P.eps[i][0] = (unsigned short)s.eps[i].x_col; P.eps[i][1] = (unsigned short)s.eps[i].s_col;
        P.eps[i][2] = (unsigned short)s.eps[i].t_col; P.eps[i][3] = 0;
    }
}

}  // namespace

cudaError_t mma_setup(const MmaPlan& m, int num_sms, DevMma& d) {
    cudaError_t e;
    d.host_prog = new MmaParamProg[3];
    fill_param_prog(m.fwd, PROG_FWD, d.host_prog[0]);
    fill_param_prog(m.fwd, PROG_INV, d.host_prog[1]);
    fill_param_prog(m.bwd, PROG_BWD, d.host_prog[2]);
    if ((e = setup_schedule(m.fwd, false, num_sms, d.fwd)) != cudaSuccess) return e;
    if ((e = setup_schedule(m.bwd, true, num_sms, d.bwd)) != cudaSuccess) return e;
    if ((e = upload(&d.pack_src, m.pack_src)) != cudaSuccess) return e;
    return upload(&d.unpack_src, m.unpack_src);
}

void mma_free(DevMma& d) {
    delete[] d.host_prog;
    d.host_prog = nullptr;
    for (DevMmaSchedule* s : {&d.fwd, &d.bwd}) { cudaFree(s->prog); cudaFree(s->eps); }
    cudaFree(d.pack_src); cudaFree(d.unpack_src);
}

int mma_weight_copies() {
    static const int n = [] { const char* e = dev_getenv("HINT_B200_MMA_WCOPIES"); const int v = e ? std::atoi(e) : kMmaWeightCopies; return v < 1 ? 1 : (v > 64 ? 64 : v); }();
    return n;
}
long long mma_copy_stride(const MmaPlan& m) { return (m.n_packed + 63) & ~63LL; }

cudaError_t mma_pack(const MmaPlan& m, const DevMma& d, const float* params, float* hi, float* lo, cudaStream_t st) {
    const long long n = m.n_packed;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
    hint_pack_mma_kernel<<<blocks, threads, 0, st>>>(d.pack_src, params, hi, lo, n, mma_weight_copies(), mma_copy_stride(m)); HINT_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t mma_launch_fwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, const float* x, const float* c,
                           const float* hi, const float* lo, float* z, float* logdet, long long B, int rev, cudaStream_t st) {
    const MSchedule& s = m.fwd;
    MmaTables T = make_tables(p, s, d.fwd, rev ? PROG_INV : PROG_FWD);
    T.wcopies = mma_weight_copies(); T.wstride = mma_copy_stride(m);
    dbg_begin(T, st);
    const MmaParamProg& P = d.host_prog[rev ? 1 : 0];
    const long long ntiles = (B + s.TM - 1) / s.TM;
    const int grid = (int)std::min<long long>(ntiles, d.fwd.max_ctas);
#define HINT_LAUNCH(TMV)                                                                                                   \
    case TMV:                                                                                                              \
        if (x3) { hint_fwd_mma_kernel<TMV, true><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, P, x, c, hi, lo, z, logdet, B, rev ? 1 : 0); HINT_LAUNCHED(); }  \
        else { hint_fwd_mma_kernel<TMV, false><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, P, x, c, hi, lo, z, logdet, B, rev ? 1 : 0); HINT_LAUNCHED(); }    \
        break;
    switch (s.TM) {
        HINT_LAUNCH(64) HINT_LAUNCH(32)
        default: return cudaErrorInvalidValue;
    }
#undef HINT_LAUNCH
    dbg_end(rev ? "inverse" : "forward", st);
    return cudaGetLastError();
}

cudaError_t mma_launch_bwd(const Plan& p, const MmaPlan& m, const DevMma& d, bool x3, int grid, const float* z, const float* c,
                           const float* hi, const float* lo, const float* dz, const float* dlogdet, float* x_rec, float* dx,
                           float* dc, float* partials, long long B, cudaStream_t st) {
    const MSchedule& s = m.bwd;
    MmaTables T = make_tables(p, s, d.bwd, PROG_BWD);
    T.wcopies = mma_weight_copies(); T.wstride = mma_copy_stride(m);
    dbg_begin(T, st);
    const MmaParamProg& P = d.host_prog[2];
    const long long np = m.n_partial;
#define HINT_LAUNCH(TMV)                                                                                                   \
    case TMV:                                                                                                              \
        if (x3) { hint_bwd_mma_kernel<TMV, true><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, P, z, c, hi, lo, dz, dlogdet, x_rec, dx, dc, partials, np, B); HINT_LAUNCHED(); }  \
        else { hint_bwd_mma_kernel<TMV, false><<<grid, kMmaThreads, s.smem_bytes, st>>>(T, P, z, c, hi, lo, dz, dlogdet, x_rec, dx, dc, partials, np, B); HINT_LAUNCHED(); }    \
        break;
    switch (s.TM) {
        HINT_LAUNCH(64) HINT_LAUNCH(32)
        default: return cudaErrorInvalidValue;
    }
#undef HINT_LAUNCH
    dbg_end("backward", st);
    return cudaGetLastError();
}

}  // namespace hint
