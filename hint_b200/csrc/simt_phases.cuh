// Phase functions of the fused-tree FP32 kernels (HINT_MODE_FP32).
//
// One CTA owns a tile of TM samples for the whole coupling tree.  All tile state lives in shared
// memory, column-major: column c, sample m at S[c*(TM+4) + m].  The x tile, the hidden activations of
// the node(s) being processed and their s/t outputs are columns; torch.split / torch.cat of
// hint.py:68,90 are pure column bookkeeping (no data movement).
//
// Every phase is a function of (tid, shared memory, descriptor tables) with no intra-phase
// cross-thread dependency; the kernels in simt_kernels.cu put a __syncthreads() between phases.
// The functions are __host__ __device__ so tests/emul can run the very same index logic on a
// CPU-only machine (thread loop instead of a CTA) against the oracle.
#pragma once
#include "plan.h"

#if defined(__CUDACC__)
#define HINT_HD __host__ __device__ __forceinline__
#else
#define HINT_HD inline
struct float4 { float x, y, z, w; };
#endif

namespace hint {

struct DevTables {
    const CG* cgs;
    const Ep* eps;
    const DwJob* dwjobs;
    const Stage* stages;
    int nstages;
    int d, dc;
    int col_x, col_d, col_one, col_zero;
    int raw_off;
    float alpha;
};

HINT_HD float4 ldg4(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}

HINT_HD void fma8x4(float (&acc)[4][8], const float4& w, const float4& a0, const float4& a1) {
    const float wv[4] = {w.x, w.y, w.z, w.w};
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(av[i], wv[j], acc[j][i]);
}

// out[TM x 4] (op)= in[TM x K] * W[K x 4] (+ bias) for the column groups [begin, end).
// Thread -> (sample group of 8, column group); consecutive threads take consecutive sample groups so
// activation loads of a quarter warp are 256 contiguous bytes and the 128-bit weight load is shared.
template <int TM>
HINT_HD void run_cgs(int tid, float* S, const CG* __restrict__ cgs, int begin, int end, const float* __restrict__ W) {
    constexpr int TMS = TM + 4, MG = TM / 8, CPI = kThreads / MG;
    const int mg = tid % MG;
    for (int g = begin + tid / MG; g < end; g += CPI) {
        const CG cg = cgs[g];
        float acc[4][8];
        if (cg.b_off >= 0) {
            const float4 b = ldg4(W + cg.b_off);
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = bv[j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
        }
        const float* w = W + cg.w_off;
        const float* a = S + cg.in0 * TMS + mg * 8;
#pragma unroll 4
        for (int k = 0; k < cg.k0; ++k) {
            const float4 wv = ldg4(w);
            const float4 a0 = *reinterpret_cast<const float4*>(a);
            const float4 a1 = *reinterpret_cast<const float4*>(a + 4);
            fma8x4(acc, wv, a0, a1);
            w += cg.ldw;
            a += TMS;
        }
        a = S + cg.in1 * TMS + mg * 8;
        for (int k = 0; k < cg.k1; ++k) {
            const float4 wv = ldg4(w);
            const float4 a0 = *reinterpret_cast<const float4*>(a);
            const float4 a1 = *reinterpret_cast<const float4*>(a + 4);
            fma8x4(acc, wv, a0, a1);
            w += cg.ldw;
            a += TMS;
        }
        float* o = S + cg.out0 * TMS + mg * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < cg.nvalid) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = acc[j][i];
                if (cg.flags & CG_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (cg.flags & (CG_MASK | CG_ACCUM)) {
                    const float4 o0 = *reinterpret_cast<const float4*>(o);
                    const float4 o1 = *reinterpret_cast<const float4*>(o + 4);
                    const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                    if (cg.flags & CG_MASK) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = ov[i] > 0.f ? v[i] : 0.f;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += ov[i];
                    }
                }
                float4 r0, r1;
                r0.x = v[0]; r0.y = v[1]; r0.z = v[2]; r0.w = v[3];
                r1.x = v[4]; r1.y = v[5]; r1.z = v[6]; r1.w = v[7];
                *reinterpret_cast<float4*>(o) = r0;
                *reinterpret_cast<float4*>(o + 4) = r1;
            }
            o += TMS;
        }
    }
}

// Coupling of hint.py:79-84 on the lower-half columns listed in eps[begin, end).
//   forward:  x_l <- exp(alpha*atan s) * x_l + t ,  J += alpha*atan s
//   inverse:  x_l <- (x_l - t) / exp(alpha*atan s), J -= alpha*atan s
// Thread -> (sample m, column slot); JP[tid] is that thread's private log-det partial.
template <int TM>
HINT_HD void coupling(int tid, float* S, const Ep* __restrict__ eps, int begin, int end, float alpha, int col_x,
                      float* JP, bool rev) {
    constexpr int TMS = TM + 4, NJG = kThreads / TM;
    const int m = tid % TM;
    float jacc = 0.f;
    for (int e = begin + tid / TM; e < end; e += NJG) {
        const Ep ep = eps[e];
        const float s = S[ep.s_col * TMS + m];
        const float t = S[ep.t_col * TMS + m];
        const float la = alpha * atanf(s);
        float* xp = S + (col_x + ep.x_col) * TMS + m;
        if (!rev) {
            *xp = expf(la) * (*xp) + t;
            jacc += la;
        } else {
            *xp = (*xp - t) / expf(la);
            jacc -= la;
        }
    }
    JP[tid] += jacc;
}

// Backward coupling: given the node OUTPUT z_l (x tile), upstream dz_l (gradient tile) and dJ:
//   x_l' = (z_l - t)/e ;  dt = dz_l ;  ds = (dz_l*(z_l - t) + dJ) * alpha/(1+s^2) ;  dx_l' = dz_l*e
// The s/t output columns are overwritten with ds/dt.
template <int TM>
HINT_HD void coupling_bwd(int tid, float* S, const Ep* __restrict__ eps, int begin, int end, float alpha, int col_x,
                          int col_d, const float* DJ) {
    constexpr int TMS = TM + 4, NJG = kThreads / TM;
    const int m = tid % TM;
    const float dj = DJ[m];
    for (int e = begin + tid / TM; e < end; e += NJG) {
        const Ep ep = eps[e];
        float* sp = S + ep.s_col * TMS + m;
        float* tp = S + ep.t_col * TMS + m;
        float* xp = S + (col_x + ep.x_col) * TMS + m;
        float* dp = S + (col_d + ep.x_col) * TMS + m;
        const float s = *sp, t = *tp;
        const float ex = expf(alpha * atanf(s));
        const float zl = *xp, dzl = *dp;
        const float r = zl - t;
        *xp = r / ex;
        *dp = dzl * ex;
        *tp = dzl;
        *sp = (dzl * r + dj) * (alpha / (1.f + s * s));
    }
}

// dW[4 rows x 4 cols per thread] = sum_m dOut[row][m] * In[col][m] over the tile, then added to this
// CTA's private partial buffer (plain read-modify-write: the region is owned by the CTA).
template <int TM>
HINT_HD void run_dw(int tid, const float* S, const DwJob* __restrict__ jobs, int jb, int je, int nitems, int nbitems,
                    float* partial, bool first, int col_zero) {
    constexpr int TMS = TM + 4;
    for (int it = tid; it < nitems; it += kThreads) {
        int lo = jb, hi = je - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].item_begin <= it) lo = mid; else hi = mid - 1;
        }
        const DwJob jd = jobs[lo];
        const int local = it - jd.item_begin;
        const int lane = local & 3;
        const int ng = (local >> 2) % jd.nN;
        const int kb = (local >> 2) / jd.nN;
        const float* np = S + (jd.n_col + 4 * ng) * TMS;
        const float* kp[4];
        int kk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            kk[j] = 16 * kb + lane + 4 * j;
            const int col = kk[j] < jd.k0 ? jd.in0 + kk[j] : (kk[j] < jd.k0 + jd.k1 ? jd.in1 + kk[j] - jd.k0 : col_zero);
            kp[j] = S + col * TMS;
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int m = 0; m < TM; m += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(np + i * TMS + m);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(kp[j] + m);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
        float* o = partial + jd.out_off + (4 * ng) * jd.ld;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float* q = o + i * jd.ld + kk[j];
                *q = first ? acc[i][j] : (*q + acc[i][j]);
            }
    }
    // bias gradients: db[row] = sum_m dOut[row][m]; one thread per row, consecutive threads -> consecutive columns
    for (int it = tid; it < nbitems; it += kThreads) {
        int lo = jb, hi = je - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].bitem_begin <= it) lo = mid; else hi = mid - 1;
        }
        const DwJob jd = jobs[lo];
        const int row = it - jd.bitem_begin;
        const float* np = S + (jd.n_col + row) * TMS;
        float acc = 0.f;
        for (int m = 0; m < TM; m += 4) {
            const float4 a = *reinterpret_cast<const float4*>(np + m);
            acc += (a.x + a.y) + (a.z + a.w);
        }
        float* q = partial + jd.b_off + row;
        *q = first ? acc : (*q + acc);
    }
}

// Row-major global tile [rows x width] (a contiguous chunk) -> columns [col_base, col_base+width).
// Rows past the end of the batch are filled with zeros.
template <int TM>
HINT_HD void load_tile(int tid, float* S, int col_base, const float* __restrict__ g, long long row0, long long B, int width) {
    constexpr int TMS = TM + 4;
    if (width == 0) return;
    const long long base = row0 * width;
    const long long rows = (B - row0) < TM ? (B - row0) : TM;
    const int nvalid = (int)(rows * width);
    const int n = TM * width;
    for (int i = tid * 4; i < n; i += kThreads * 4) {
        float v[4];
        if (i + 3 < nvalid) {
            const float4 q = ldg4(g + base + i);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (i + e < nvalid) ? g[base + i + e] : 0.f;
        }
        int m = i / width, j = i - m * width;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (i + e < n) S[(col_base + j) * TMS + m] = v[e];
            if (++j == width) { j = 0; ++m; }
        }
    }
}

template <int TM>
HINT_HD void store_tile(int tid, const float* S, int col_base, float* __restrict__ g, long long row0, long long B, int width) {
    constexpr int TMS = TM + 4;
    if (width == 0) return;
    const long long base = row0 * width;
    const long long rows = (B - row0) < TM ? (B - row0) : TM;
    const int nvalid = (int)(rows * width);
    for (int i = tid * 4; i < nvalid; i += kThreads * 4) {
        float v[4];
        int m = i / width, j = i - m * width;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = (i + e < nvalid) ? S[(col_base + j) * TMS + m] : 0.f;
            if (++j == width) { j = 0; ++m; }
        }
        if (i + 3 < nvalid) {
            float4 q; q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3];
            *reinterpret_cast<float4*>(g + base + i) = q;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < nvalid) g[base + i + e] = v[e];
        }
    }
}

template <int TM>
HINT_HD void fill_col(int tid, float* S, int col, float v) {
    constexpr int TMS = TM + 4;
    for (int m = tid; m < TM; m += kThreads) S[col * TMS + m] = v;
}

}  // namespace hint
