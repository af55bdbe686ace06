// Launch glue of the tcgen05 training kernel.
#include "tc3_launch.h"
#include "launch_count.h"

#include <algorithm>
#include <cstring>
#include <vector>

#include "tc3_kernels.cuh"

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
}  // namespace

cudaError_t tc3_setup(const T3Plan& t, int num_sms, DevTc3& d) {
    static_assert(sizeof(T3Prog) < 32000, "the program must fit the kernel-parameter space");
    if ((int)t.mmas.size() > kT3MaxMma) return cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = upload(&d.epis, t.epis)) != cudaSuccess) return e;
    if ((e = upload(&d.chunks, t.chunks)) != cudaSuccess) return e;
    if ((e = upload(&d.tab16, t.tab16)) != cudaSuccess) return e;
    if ((e = upload(&d.pack_src, t.pack_src)) != cudaSuccess) return e;
    std::vector<int32_t> dst((size_t)t.n_partial, -1);
    if (t.kind == T3K_BACKWARD)
        for (size_t i = 0; i < t.unpack_src.size(); ++i) dst[(size_t)t.unpack_src[i]] = (int32_t)i | (t.unpack_q4[i] ? (1 << 30) : 0);
    if ((e = upload(&d.part_dst, dst)) != cudaSuccess) return e;
    T3Prog* P = new T3Prog();
    std::memset(P, 0, sizeof(T3Prog));
    P->n_mma = (int)t.mmas.size(); P->n_epi = (int)t.epis.size(); P->n_chunks = (int)t.chunks.size(); P->n_signals = t.n_mma_signals;
    P->n_slots = t.n_slots; P->slot_bytes = t.slot_bytes;
    P->sm_bars = t.sm_bars; P->sm_tab16 = t.sm_tab16; P->sm_epis = t.sm_epis; P->sm_xs = t.sm_xs; P->sm_gs = t.sm_gs; P->sm_os = t.sm_os; P->sm_red = t.sm_red; P->sm_ring = t.sm_ring;
    for (int i = 0; i < kT3Imgs; ++i) { P->sm_img[i] = t.sm_img[i]; P->img_rows[i] = t.img_rows[i]; }
    P->xp = t.xp; P->op = t.op; P->d = t.d; P->dc = t.dc; P->n_tab16 = (int)t.tab16.size();
    P->alpha = t.alpha;
    P->kind = t.kind;
    P->epis = d.epis; P->chunks = d.chunks; P->tab16 = d.tab16;
    for (size_t i = 0; i < t.mmas.size(); ++i) P->mmas[i] = t3_pack_mma(t.mmas[i]);
    d.prog = P;
    d.num_sms = num_sms;
    // the attribute is per function, and several plans (backward, forward, inverse; every block) share the kernel: always the maximum
    return cudaFuncSetAttribute((const void*)hint_tc3_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
}

void tc3_free(DevTc3& d) {
    cudaFree(d.epis); cudaFree(d.chunks); cudaFree(d.tab16); cudaFree(d.pack_src); cudaFree(d.part_dst);
    delete d.prog;
    d = DevTc3();
}

cudaError_t tc3_pack(const T3Plan& t, const DevTc3& d, const float* params, float* packed, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (int)std::min<long long>((t.n_packed + threads - 1) / threads, 148 * 8);
    hint_tc3_pack_kernel<<<blocks, threads, 0, st>>>(d.pack_src, params, packed, t.n_packed); HINT_LAUNCHED();
    return cudaGetLastError();
}

int tc3_bwd_ctas(const DevTc3& d, long long B) {
    return (int)std::min<long long>((B + 127) / 128, d.num_sms);
}

cudaError_t tc3_launch_bwd(const T3Plan& t, const DevTc3& d, int grid, const float* z, const float* cond, const float* packed,
                           const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc, float* partials,
                           float* dparams, long long B, cudaStream_t st, long long* prof, float nll_scale) {
    if (nll_scale != 0.f) {   // the program travels by value: a per-launch field is a host-side copy (the stored plan stays immutable)
        T3Prog* P = new T3Prog(*d.prog);
        P->nll_scale = nll_scale;
        hint_tc3_bwd_kernel<<<grid, kT3Threads, t.smem_bytes, st>>>(*P, z, cond, packed, dz, dlogdet, x_rec, dx, dc, partials,
                                                                    (long long)t.n_partial, B, nullptr, prof); HINT_LAUNCHED();
        delete P;
    } else {
        hint_tc3_bwd_kernel<<<grid, kT3Threads, t.smem_bytes, st>>>(*d.prog, z, cond, packed, dz, dlogdet, x_rec, dx, dc, partials,
                                                                    (long long)t.n_partial, B, nullptr, prof); HINT_LAUNCHED();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((t.n_partial + threads - 1) / threads, 148 * 8);
    hint_tc3_reduce_kernel<<<blocks, threads, 0, st>>>(d.part_dst, partials, grid, (long long)t.n_partial, dparams); HINT_LAUNCHED();
    return cudaGetLastError();
}


cudaError_t tc3_launch_transport(const T3Plan& t, const DevTc3& d, const float* x, const float* cond, const float* packed, float* z,
                                 float* logdet, long long B, cudaStream_t st) {
    const int grid = tc3_bwd_ctas(d, B);
    hint_tc3_bwd_kernel<<<grid, kT3Threads, t.smem_bytes, st>>>(*d.prog, x, cond, packed, nullptr, nullptr, logdet, z, nullptr, nullptr, 0, B,
                                                                nullptr, nullptr); HINT_LAUNCHED();
    return cudaGetLastError();
}

// Developer aid (tests/cuda/dbg_tc3.py): run ONE tile, stop after `n_epi_limit` epilogue steps and dump TMEM + shared memory.
cudaError_t tc3_debug_run(const T3Plan& t, const DevTc3& d, int n_epi_limit, const float* z, const float* cond, const float* packed,
                          const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc, float* partials, long long B,
                          float* dump, cudaStream_t st) {
    T3Prog P = *d.prog;
    n_epi_limit = std::min(n_epi_limit, (int)t.epis.size());
    int nm = 0, nch = 0;
    while (nm < (int)t.mmas.size() && t.mmas[nm].wait_epi < n_epi_limit) {
        if (!(t.mmas[nm].flags & T3M_SS) && (t.mmas[nm].flags & T3M_NEWCHUNK)) ++nch;
        ++nm;
    }
    // a slab must be consumed completely: extend to the record that releases the last chunk
    while (nm < (int)t.mmas.size() && nm > 0 && !(t.mmas[nm - 1].flags & T3M_SS) && !(t.mmas[nm - 1].flags & T3M_ENDCHUNK)) ++nm;
    P.n_mma = nm; P.n_epi = n_epi_limit; P.n_chunks = nch;
    hint_tc3_bwd_kernel<<<1, kT3Threads, t.smem_bytes, st>>>(P, z, cond, packed, dz, dlogdet, x_rec, dx, dc, partials,
                                                             (long long)t.n_partial, std::min<long long>(B, 128), dump, nullptr); HINT_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace hint
