// Fused-tree FP32 kernels: one persistent CTA per (SM x occupancy), each looping over tiles of TM
// samples and walking the whole coupling tree for its tile out of shared memory.
//   forward / inverse : hint.py:62-101   (stages deepest-first for rev=0, root-first for rev=1)
//   backward          : the autograd tape of the same lines, restated as a root-first sweep that
//                       inverts the block output on the fly (no stored activations).
#pragma once
#include "simt_phases.cuh"

namespace hint {

// A phase = the same statement executed by every thread of the CTA, followed by a barrier.  On the
// device `tid` is threadIdx.x; in the host emulation (tests/emul) the CTA is a plain loop.
#if defined(__CUDA_ARCH__)
#define HINT_PHASE(...) { const int tid = threadIdx.x; __VA_ARGS__; } __syncthreads();
#else
// host emulation: visit the thread ids in a permuted order (stride coprime to 256) so an intra-phase dependency
// between threads cannot hide behind sequential execution
#ifndef HINT_EMUL_STRIDE
#define HINT_EMUL_STRIDE 1
#endif
#define HINT_PHASE(...) for (int t_ = 0; t_ < kThreads; ++t_) { const int tid = (t_ * HINT_EMUL_STRIDE + 11) % kThreads; __VA_ARGS__; }
#endif

// One tile of TM samples through the whole tree: transport + log-det.
template <int TM>
HINT_HD void fwd_tile(const DevTables& T, float* S, const float* __restrict__ x, const float* __restrict__ c,
                      const float* __restrict__ W, float* __restrict__ z, float* __restrict__ logdet, long long B,
                      int rev, long long row0) {
    float* JP = S + T.raw_off;
    HINT_PHASE(load_tile<TM>(tid, S, T.col_x, x, row0, B, T.d);
               load_tile<TM>(tid, S, T.col_x + T.d, c, row0, B, T.dc);
               JP[tid] = 0.f)
    for (int i = 0; i < T.nstages; ++i) {
        const Stage& st = T.stages[rev ? i : T.nstages - 1 - i];
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[0], st.cg_begin[1], W))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[1], st.cg_begin[2], W))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[2], st.cg_begin[3], W))
        HINT_PHASE(coupling<TM>(tid, S, T.eps, st.ep_begin, st.ep_end, T.alpha, T.col_x, JP, rev != 0))
    }
    HINT_PHASE(store_tile<TM>(tid, S, T.col_x, z, row0, B, T.d);
               if (tid < TM && row0 + tid < B) {
                   float j = 0.f;
                   for (int q = 0; q < kThreads / TM; ++q) j += JP[q * TM + tid];
                   logdet[row0 + tid] = j;
               })
}

// One tile through the backward sweep (root first): recompute, invert, emit gradients.
template <int TM>
HINT_HD void bwd_tile(const DevTables& T, float* S, const float* __restrict__ z, const float* __restrict__ c,
                      const float* __restrict__ W, const float* __restrict__ dz, const float* __restrict__ dlogdet,
                      float* __restrict__ x_rec, float* __restrict__ dx, float* __restrict__ dc, float* partial,
                      bool first, long long B, long long row0) {
    float* DJ = S + T.raw_off;
    HINT_PHASE(load_tile<TM>(tid, S, T.col_x, z, row0, B, T.d);
               load_tile<TM>(tid, S, T.col_x + T.d, c, row0, B, T.dc);
               load_tile<TM>(tid, S, T.col_d, dz, row0, B, T.d);
               for (int j = 0; j < T.dc; ++j) fill_col<TM>(tid, S, T.col_d + T.d + j, 0.f);
               fill_col<TM>(tid, S, T.col_zero, 0.f);
               if (tid < TM) DJ[tid] = (row0 + tid < B) ? dlogdet[row0 + tid] : 0.f)
    for (int i = 0; i < T.nstages; ++i) {
        const Stage& st = T.stages[i];
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[0], st.cg_begin[1], W))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[1], st.cg_begin[2], W))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[2], st.cg_begin[3], W))
        HINT_PHASE(coupling_bwd<TM>(tid, S, T.eps, st.ep_begin, st.ep_end, T.alpha, T.col_x, T.col_d, DJ))
        HINT_PHASE(run_dw<TM>(tid, S, T.dwjobs, st.dw_begin[0], st.dw_begin[1], st.dw_items[0], st.dw_bitems[0], partial, first, T.col_zero))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[3], st.cg_begin[4], W))  // dH2, in place over H2
        HINT_PHASE(run_dw<TM>(tid, S, T.dwjobs, st.dw_begin[1], st.dw_begin[2], st.dw_items[1], st.dw_bitems[1], partial, first, T.col_zero))
        HINT_PHASE(run_cgs<TM>(tid, S, T.cgs, st.cg_begin[4], st.cg_begin[5], W))  // dH1, in place over H1
        HINT_PHASE(run_dw<TM>(tid, S, T.dwjobs, st.dw_begin[2], st.dw_begin[3], st.dw_items[2], st.dw_bitems[2], partial, first, T.col_zero);
                   run_cgs<TM>(tid, S, T.cgs, st.cg_begin[5], st.cg_begin[7], W))   // dx_upper +=, dc +=
    }
    HINT_PHASE(if (x_rec) store_tile<TM>(tid, S, T.col_x, x_rec, row0, B, T.d);
               store_tile<TM>(tid, S, T.col_d, dx, row0, B, T.d);
               if (dc) store_tile<TM>(tid, S, T.col_d + T.d, dc, row0, B, T.dc))
}

#if defined(__CUDACC__)
template <int TM>
__global__ void __launch_bounds__(kThreads)
hint_fwd_fp32_kernel(DevTables T, const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ W,
                     float* __restrict__ z, float* __restrict__ logdet, long long B, int rev) {
    extern __shared__ float4 smem4[];
    float* S = reinterpret_cast<float*>(smem4);
    const long long ntiles = (B + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        fwd_tile<TM>(T, S, x, c, W, z, logdet, B, rev, tile * TM);
}

template <int TM>
__global__ void __launch_bounds__(kThreads)
hint_bwd_fp32_kernel(DevTables T, const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ W,
                     const float* __restrict__ dz, const float* __restrict__ dlogdet, float* __restrict__ x_rec,
                     float* __restrict__ dx, float* __restrict__ dc, float* __restrict__ partials, long long n_partial,
                     long long B) {
    extern __shared__ float4 smem4[];
    float* S = reinterpret_cast<float*>(smem4);
    float* partial = partials + (long long)blockIdx.x * n_partial;
    const long long ntiles = (B + TM - 1) / TM;
    bool first = true;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        bwd_tile<TM>(T, S, z, c, W, dz, dlogdet, x_rec, dx, dc, partial, first, B, tile * TM);
        first = false;
    }
}

__global__ void hint_pack_kernel(const int* __restrict__ src, const float* __restrict__ params, float* __restrict__ packed,
                                 long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = src[i];
        packed[i] = s < 0 ? 0.f : params[s];
    }
}

// dparams[i] = sum over the CTAs' partial buffers, in fixed CTA order (deterministic, no atomics)
__global__ void hint_reduce_unpack_kernel(const int* __restrict__ src, const float* __restrict__ partials, int nctas,
                                          long long n_partial, float* __restrict__ dparams, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* p = partials + src[i];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int q = 0;
        for (; q + 3 < nctas; q += 4) {
            a0 += p[(long long)q * n_partial];
            a1 += p[(long long)(q + 1) * n_partial];
            a2 += p[(long long)(q + 2) * n_partial];
            a3 += p[(long long)(q + 3) * n_partial];
        }
        for (; q < nctas; ++q) a0 += p[(long long)q * n_partial];
        dparams[i] = (a0 + a1) + (a2 + a3);
    }
}

#endif  // __CUDACC__

}  // namespace hint
