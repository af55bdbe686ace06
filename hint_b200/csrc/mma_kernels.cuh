// Warp-MMA fused-tree kernels (HINT_MODE_TF32 / HINT_MODE_TF32X3): forward / inverse transport + log-det and the
// memory-free backward of one HINT coupling block, subnet GEMMs on mma.sync.m16n8k8 tf32 with fp32 accumulation.
//   forward / inverse : hint.py:62-101      backward : the autograd tape of the same lines (SURVEY 8a formulas)
// Schedule, layouts and the reason this path exists next to the tcgen05 one: plan_mma.h.
//
// The code is written against four primitives (CTA barrier, warp MMA, read-only load, lane id) so that the very same
// functions run under tests/emul/emul_mma.cpp, where a CTA is 256 fibers and mma_tf32 is a lane-exchange that
// truncates its operands to 10 mantissa bits like the tensor core does.
#pragma once
#include <cstdint>
#include <cstring>

#include "plan_mma.h"

#if defined(__CUDACC__)
#define HINT_DEV __device__ __forceinline__
#define HINT_DEV_CALL __device__ __noinline__   // real calls: each task shape gets its own register allocation
#else
#define HINT_DEV_CALL inline
#include <cmath>
#define HINT_DEV inline
namespace hint { namespace emu {
void cta_sync();
void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
} }
#ifndef __restrict__
#define __restrict__
#endif
#endif

#ifndef HINT_MMA_EXP
#define HINT_MMA_EXP 0   // developer experiments (timing only, wrong results): 2 no weight loads, 8 no MMA, 16 no A loads
#endif

namespace hint {

struct MmaTables {
    const WOp* prog;            // op streams (plan_mma.h) in global memory ...
    const Ep* eps;
    int in_param;               // ... or, when they fit, in the MmaParamProg kernel parameter (then `begin` indexes P.ops)
    int begin[kMmaWarps];       // first op of each warp's stream for this launch (forward / inverse / backward program)
    int d, dc;
    int col_x, col_d, col_one, col_zero;
    int raw_off;
    float alpha;
    int wcopies;                // the packed operands are replicated `wcopies` times, `wstride` floats apart: CTA b reads copy
    long long wstride;          // b % wcopies, so the 2*SMs CTAs running the same program do not hammer the same L2 lines
    long long* dbg;             // optional: clock64 after every barrier of CTA 0's second tile ([0] = count) - HINT_B200_MMA_DEBUG
};

// ---- primitives ----------------------------------------------------------------------------------------------------
HINT_DEV void m_dbg_stamp(long long* dbg, int tid, int bid) {
#if defined(__CUDA_ARCH__)
    if (dbg != nullptr && tid == 0 && bid == 0) {
        const long long n = dbg[0];
        if (n < 1000) { dbg[1 + n] = clock64(); dbg[0] = n + 1; }
    }
#endif
}

HINT_DEV void m_cta_sync() {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#elif !defined(__CUDACC__)
    emu::cta_sync();
#endif
}

HINT_DEV void m_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if defined(__CUDA_ARCH__) && (HINT_MMA_EXP & 8)
    c[0] += __uint_as_float(a[0]) * __uint_as_float(b0); c[3] += __uint_as_float(a[3]) * __uint_as_float(b1);
#elif defined(__CUDA_ARCH__)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#elif !defined(__CUDACC__)
    emu::mma_tf32(c, a, b0, b1);
#endif
}

HINT_DEV uint32_t m_bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v);
#else
    uint32_t u; std::memcpy(&u, &v, 4); return u;
#endif
}
HINT_DEV float m_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float v; std::memcpy(&v, &u, 4); return v;
#endif
}
// round to nearest (ties away from zero) to tf32: the tensor core ignores the low 13 mantissa bits, so adding half an
// ulp of the 10-bit mantissa to the magnitude is all the rounding needs (cvt.rna.tf32 runs at 16 lanes/clk/SM)
HINT_DEV uint32_t m_rna(float v) { return m_bits(v) + 0x1000u; }
HINT_DEV float m_trunc_tf32(uint32_t u) { return m_float(u & 0xFFFFE000u); }

HINT_DEV void m_ld2(const float* __restrict__ p, float& a, float& b) {
#if defined(__CUDA_ARCH__)
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
// B fragments stream from L2 and are read once per task: do not let them evict the op records from L1
HINT_DEV void m_ld2_stream(const float* __restrict__ p, float& a, float& b) {
#if defined(__CUDA_ARCH__) && (HINT_MMA_EXP & 2)
    a = 0.01f; b = -0.01f;
#elif defined(__CUDA_ARCH__) && (HINT_MMA_EXP & 1)
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    a = v.x; b = v.y;
#elif defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
// partial[0..1] += {a, b}: fire-and-forget reduction at L2 (the address is owned by this thread: order = program order)
HINT_DEV void m_red2(float* p, float a, float b) {
#if defined(__CUDA_ARCH__)
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
#else
    p[0] += a; p[1] += b;
#endif
}
// exp / atan of the soft clamp (hint.py:56-60).  exp through ex2 (|x| <= clamp*0.636*pi/2: relative error ~2e-7);
// atan by reciprocal range reduction + degree-7 minimax polynomial in x^2 (max abs error 1.7e-7).
HINT_DEV float m_exp(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
#else
    return exp2f(x * 1.4426950408889634f);
#endif
}
HINT_DEV float m_rcp(float a) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.f / a;
#endif
}
HINT_DEV float m_atan(float x) {
    const float a = fabsf(x);
    const bool inv = a > 1.f;
    const float r = inv ? m_rcp(a) : a;
    const float t = r * r;
    float p = -0.004780583083629608f;
    p = fmaf(p, t, 0.024557599797844887f);
    p = fmaf(p, t, -0.05990542098879814f);
    p = fmaf(p, t, 0.09942812472581863f);
    p = fmaf(p, t, -0.1402944177389145f);
    p = fmaf(p, t, 0.199713796377182f);
    p = fmaf(p, t, -0.3333209455013275f);
    p = fmaf(p, t, 0.9999999403953552f);
    float y = p * r;
    if (inv) y = 1.5707963267948966f - y;
    return copysignf(y, x);
}

// Operand packing (pack_src encoding: s >= 0 weight params[s], rounded to tf32; s == -1 zero; s <= -2 bias params[-s-2],
// kept exact because it initialises the fp32 accumulator).  hi = rna_tf32(w), lo = rna_tf32(w - hi) (3xTF32 split).
HINT_DEV void m_pack_elem(int s, const float* __restrict__ params, float& hi, float& lo) {
    lo = 0.f;
    if (s == -1) { hi = 0.f; return; }
    if (s <= -2) { hi = params[-s - 2]; return; }
    const float w = params[s];
    hi = m_trunc_tf32(m_rna(w));
    lo = m_trunc_tf32(m_rna(w - hi));
}

// round-to-nearest (ties away) to tf32 as a float (what the TF32 mode stores for every activation that is a GEMM operand,
// so operand loads need no conversion)
HINT_DEV float m_round_tf32(float v) { return m_trunc_tf32(m_rna(v)); }

// ---- op records decoded into registers -------------------------------------------------------------------------------
struct GemmOp { int w_off, b_off, ks_stride, ksteps, in_off, k0, in1_off, k1, out_off, zero_off, nvalid, flags; };
struct DwOp { int out_off, a_off, N, n0, kf0, in_off, k0, in1_off, k1, ld, nstore, one_off, zero_off; };
HINT_DEV int m_lo16(int w) { return w & 0xffff; }
HINT_DEV int m_hi16(int w) { return (int)((unsigned)w >> 16); }
HINT_DEV GemmOp m_decode_gemm(const int (&r)[8]) {
    GemmOp o;
    o.w_off = r[0]; o.b_off = r[1];
    o.ks_stride = m_lo16(r[2]); o.ksteps = m_hi16(r[2]);
    o.in_off = m_lo16(r[3]); o.k0 = m_hi16(r[3]);
    o.in1_off = m_lo16(r[4]); o.k1 = m_hi16(r[4]);
    o.out_off = m_lo16(r[5]); o.zero_off = m_hi16(r[5]);
    o.nvalid = m_lo16(r[6]);
    o.flags = r[7] & 0xff;
    return o;
}
HINT_DEV DwOp m_decode_dw(const int (&r)[8]) {
    DwOp o;
    o.out_off = r[0];
    o.a_off = m_lo16(r[1]); o.N = m_hi16(r[1]);
    o.n0 = m_lo16(r[2]); o.kf0 = m_hi16(r[2]);
    o.in_off = m_lo16(r[3]); o.k0 = m_hi16(r[3]);
    o.in1_off = m_lo16(r[4]); o.k1 = m_hi16(r[4]);
    o.ld = m_lo16(r[5]); o.nstore = m_hi16(r[5]);
    o.one_off = m_lo16(r[6]); o.zero_off = m_hi16(r[6]);
    return o;
}
HINT_DEV int m_op_type(const int (&r)[8]) { return (int)((unsigned)r[7] >> 24); }
HINT_DEV int m_op_flags(const int (&r)[8]) { return r[7] & 0xff; }
HINT_DEV int m_gemm_nt(const int (&r)[8]) { return (r[6] >> 24) & 0xff; }
HINT_DEV int m_dw_mt(const int (&r)[8]) { return (r[7] >> 8) & 0xff; }
HINT_DEV int m_dw_nt(const int (&r)[8]) { return (r[7] >> 16) & 0xff; }
HINT_DEV void m_prefetch_l1(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}

// ---- forward-type GEMM task ------------------------------------------------------------------------------------------
// out[TM rows, 8*nt] (op)= in[TM rows, K] * B (+ bias).  A fragments from the column-major tile with the (2t, 2t+1) k-slot
// permutation (conflict-free at pitch TM+4); B fragments stream from the packed operand buffer (hi parts in W, low parts
// in Wlo when X3) through a register ring that always holds the next kPF k-steps.  The ring of a task is filled by
// m_issue_b, which the interpreter calls as early as possible: right after the PREVIOUS task's k-loop (before its epilogue
// and before the phase barrier), so the L2 latency of the first fragments is off the critical path.
// MT_FAST = one input segment of stored (already tf32-rounded) activations, K a multiple of 8: pure pointer bumps, no
// conversion.  Otherwise (layer 1) the task reads the exact x / condition columns and rounds on load.
template <bool X3, int PF>
struct BRing {
    float b[PF][kNC][2];
    float l[X3 ? PF : 1][kNC][2];
    float bias[kNC][2];
};
// ring depth (k-steps of B fragments in flight): the accumulators of a TM = 32 tile need half the registers of a TM = 64 one
template <int TM> struct RingDepth { static constexpr int value = TM >= 64 ? kPF : 2 * kPF; };

template <bool X3, int NT, int PF>
HINT_DEV void m_issue_b(const GemmOp& tk, const float* __restrict__ W, const float* __restrict__ Wlo, int lane, BRing<X3, PF>& R) {
    const int t = lane & 3;
    const float* wp = W + tk.w_off + 2 * lane;
    const float* wl = X3 ? Wlo + tk.w_off + 2 * lane : nullptr;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        R.bias[j][0] = 0.f; R.bias[j][1] = 0.f;
        if (tk.b_off >= 0) m_ld2(W + tk.b_off + 8 * j + 2 * t, R.bias[j][0], R.bias[j][1]);
    }
#pragma unroll
    for (int s = 0; s < PF; ++s) {
        if (s < tk.ksteps) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                m_ld2_stream(wp + j * 64, R.b[s][j][0], R.b[s][j][1]);
                if (X3) m_ld2_stream(wl + j * 64, R.l[s][j][0], R.l[s][j][1]);
            }
        }
        wp += tk.ks_stride;
        if (X3) wl += tk.ks_stride;
    }
}
template <bool X3, int PF>
HINT_DEV void m_issue_b_any(const int (&r)[8], const float* __restrict__ W, const float* __restrict__ Wlo, int lane, BRing<X3, PF>& R) {
    const GemmOp tk = m_decode_gemm(r);
    const int nt = m_gemm_nt(r);
    if (nt == 3) m_issue_b<X3, 3, PF>(tk, W, Wlo, lane, R);
    else if (nt == 2) m_issue_b<X3, 2, PF>(tk, W, Wlo, lane, R);
    else m_issue_b<X3, 1, PF>(tk, W, Wlo, lane, R);
}

// Row <-> sample mapping of a forward-type task (TM = 64): MMA row (m-tile i, half hh, group g) is sample
//   32*(jj/4) + 4*g + jj%4   with jj = 2*i + hh,
// i.e. every lane owns two runs of 4 consecutive samples, so one column of A fragments (and one column of C fragments in
// the epilogue) is two 128-bit shared-memory accesses instead of eight 32-bit ones; at pitch TM+4 the quarter-warps of those
// accesses (g in {2q, 2q+1}, t in 0..3 -> word 8t + 4g) are bank-conflict free.  Samples are independent, so the mapping is
// private to these two functions (the weight-gradient tasks contract over samples: any order).
template <int TM>
HINT_DEV void m_ld_col(const float* p, float (&v)[TM / 8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int q = 0; q < TM / 32; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(p + 32 * q);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
#else
    for (int q = 0; q < TM / 32; ++q)
        for (int e = 0; e < 4; ++e) v[4 * q + e] = p[32 * q + e];
#endif
}
template <int TM>
HINT_DEV void m_st_col(float* p, const float (&v)[TM / 8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int q = 0; q < TM / 32; ++q) *reinterpret_cast<float4*>(p + 32 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#else
    for (int q = 0; q < TM / 32; ++q)
        for (int e = 0; e < 4; ++e) p[32 * q + e] = v[4 * q + e];
#endif
}

// k-loop of one task; the ring holds k-steps [0, kPF) on entry and is dead on exit
template <int TM, bool X3, int NT>
HINT_DEV void m_gemm_kloop(const GemmOp& tk, const float* S, const float* __restrict__ W, const float* __restrict__ Wlo,
                           int lane, BRing<X3, RingDepth<TM>::value>& R, float (&acc)[TM / 16][kNC][4]) {
    constexpr int TMS = TM + 4, MT = TM / 16, PF = RingDepth<TM>::value;
    static_assert(TM % 32 == 0, "the vectorised row mapping needs whole runs of 32 samples");
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            acc[i][j][0] = R.bias[j][0]; acc[i][j][1] = R.bias[j][1]; acc[i][j][2] = R.bias[j][0]; acc[i][j][3] = R.bias[j][1];
        }
    const bool fast = tk.flags & MT_FAST;
    const float* wp = W + tk.w_off + 2 * lane + PF * tk.ks_stride;
    const float* wl = X3 ? Wlo + tk.w_off + 2 * lane + PF * tk.ks_stride : nullptr;
    const int K = tk.k0 + tk.k1;
    const float* pa = S + tk.in_off + 2 * t * TMS + 4 * g;
    for (int ks = 0; ks < tk.ksteps; ++ks) {
        const float* p0 = pa;
        const float* p1 = pa + TMS;
        pa += 8 * TMS;
        if (!fast) {
            const int f0 = 8 * ks + 2 * t, f1 = f0 + 1;
            const int o0 = f0 < tk.k0 ? tk.in_off + f0 * TMS : (f0 < K ? tk.in1_off + (f0 - tk.k0) * TMS : tk.zero_off);
            const int o1 = f1 < tk.k0 ? tk.in_off + f1 * TMS : (f1 < K ? tk.in1_off + (f1 - tk.k0) * TMS : tk.zero_off);
            p0 = S + o0 + 4 * g; p1 = S + o1 + 4 * g;
        }
        float v0[TM / 8], v1[TM / 8];   // column 2t / 2t+1 of this k-step, the lane's 2*MT rows (jj = 2i + hh)
        m_ld_col<TM>(p0, v0);
        m_ld_col<TM>(p1, v1);
        uint32_t a[MT][4], al[X3 ? MT : 1][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const float v[4] = {v0[2 * i], v0[2 * i + 1], v1[2 * i], v1[2 * i + 1]};   // (g, 2t) (g+8, 2t) (g, 2t+1) (g+8, 2t+1)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (X3) {
                    a[i][e] = m_rna(v[e]);
                    al[i][e] = m_rna(v[e] - m_trunc_tf32(a[i][e]));
                } else {
                    a[i][e] = m_bits(v[e]);
                }
            }
        }
        if (!X3 && !fast) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) a[i][e] += 0x1000u;
        }
        float b[NT][2], bl[X3 ? NT : 1][2];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            b[j][0] = R.b[0][j][0]; b[j][1] = R.b[0][j][1];
            if (X3) { bl[j][0] = R.l[0][j][0]; bl[j][1] = R.l[0][j][1]; }
#pragma unroll
            for (int s = 0; s + 1 < PF; ++s) {
                R.b[s][j][0] = R.b[s + 1][j][0]; R.b[s][j][1] = R.b[s + 1][j][1];
                if (X3) { R.l[s][j][0] = R.l[s + 1][j][0]; R.l[s][j][1] = R.l[s + 1][j][1]; }
            }
        }
        if (ks + PF < tk.ksteps) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                m_ld2_stream(wp + j * 64, R.b[PF - 1][j][0], R.b[PF - 1][j][1]);
                if (X3) m_ld2_stream(wl + j * 64, R.l[PF - 1][j][0], R.l[PF - 1][j][1]);
            }
        }
        wp += tk.ks_stride;
        if (X3) wl += tk.ks_stride;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                if (X3) {
                    m_mma(acc[i][j], al[i], m_bits(b[j][0]), m_bits(b[j][1]));
                    m_mma(acc[i][j], a[i], m_bits(bl[j][0]), m_bits(bl[j][1]));
                }
                m_mma(acc[i][j], a[i], m_bits(b[j][0]), m_bits(b[j][1]));
            }
        }
    }
}

// epilogue: C fragment (row g / g+8, col 2t / 2t+1) -> column-major tile, one column of the lane's rows = two 128-bit stores.
// Three straight-line variants chosen by a warp-uniform branch: write-only (bias+ReLU or plain), mask (dH = G * [h > 0],
// in place over h), accumulate (dx_upper += ...).  Stored GEMM operands are rounded to tf32 here (not in X3 mode).
template <int TM, bool X3, int NT>
HINT_DEV void m_gemm_epilogue(const GemmOp& tk, float* S, int lane, const float (&acc)[TM / 16][kNC][4]) {
    constexpr int TMS = TM + 4, MT = TM / 16;
    const int g = lane >> 2, t = lane & 3;
    float* q = S + tk.out_off + 2 * t * TMS + 4 * g;
    const bool rnd = !X3 && (tk.flags & (MT_RELU | MT_MASK));
    const uint32_t radd = rnd ? 0x1000u : 0u, rmask = rnd ? 0xFFFFE000u : 0xFFFFFFFFu;
    const int mode = (tk.flags & MT_MASK) ? 1 : ((tk.flags & MT_ACCUM) ? 2 : 0);
    const float lo = (tk.flags & MT_RELU) ? 0.f : -3.0e38f;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (8 * j + 2 * t + c < tk.nvalid) {
                float* o = q + (8 * j + c) * TMS;
                float v[TM / 8];
#pragma unroll
                for (int i = 0; i < MT; ++i) { v[2 * i] = acc[i][j][c]; v[2 * i + 1] = acc[i][j][2 + c]; }
                if (mode == 0) {
#pragma unroll
                    for (int e = 0; e < TM / 8; ++e) v[e] = m_float((m_bits(fmaxf(v[e], lo)) + radd) & rmask);
                } else {
                    float old[TM / 8];
                    m_ld_col<TM>(o, old);
                    if (mode == 1) {
#pragma unroll
                        for (int e = 0; e < TM / 8; ++e) v[e] = m_float((m_bits(old[e] > 0.f ? v[e] : 0.f) + radd) & rmask);
                    } else {
#pragma unroll
                        for (int e = 0; e < TM / 8; ++e) v[e] += old[e];
                    }
                }
                m_st_col<TM>(o, v);
            }
}

template <int TM, bool X3, int NT>
HINT_DEV void m_gemm_task(const int (&cur)[8], const int (&nxt)[8], bool sync, bool& ring_valid, float* S,
                          const float* __restrict__ W, const float* __restrict__ Wlo, int lane, BRing<X3, RingDepth<TM>::value>& R,
                          long long* dbg, int tid) {
    const GemmOp tk = m_decode_gemm(cur);
#if HINT_MMA_EXP & 32
    m_dbg_stamp(dbg, tid, 0);
#endif
    if (!ring_valid) m_issue_b<X3, NT, RingDepth<TM>::value>(tk, W, Wlo, lane, R);
    if (sync) { m_cta_sync(); m_dbg_stamp(dbg, tid, 0); }
    float acc[TM / 16][kNC][4];
    m_gemm_kloop<TM, X3, NT>(tk, S, W, Wlo, lane, R, acc);
#if HINT_MMA_EXP & 32
    m_dbg_stamp(dbg, tid, 0);
#endif
    ring_valid = m_op_type(nxt) == OP_GEMM;
    if (ring_valid) m_issue_b_any<X3, RingDepth<TM>::value>(nxt, W, Wlo, lane, R);      // next task's first fragments fly during this epilogue (+ barrier)
    m_gemm_epilogue<TM, X3, NT>(tk, S, lane, acc);
#if HINT_MMA_EXP & 32
    m_dbg_stamp(dbg, tid, 0);
#endif
}

// ---- weight-gradient task ----------------------------------------------------------------------------------------------
// dW[rows n0.., in-features kf0..] (+)= sum over the tile's samples of dOut[sample][row] * In[sample][feature]
// (M = out rows, N = in-features, K = samples).  Both operands come from the column-major tile; the partial buffer is
// private to the CTA and each element is owned by exactly one thread -> store on the first tile, fire-and-forget red after.
template <int TM, bool X3, int MT, int NT>
HINT_DEV_CALL void m_dw(int r0, int r1, int r2, int r3, int r4, int r5, int r6, int r7, const float* S,
                        float* __restrict__ partial, bool first, int lane) {
    constexpr int TMS = TM + 4;
    const int rr[8] = {r0, r1, r2, r3, r4, r5, r6, r7};
    const DwOp tk = m_decode_dw(rr);
    const int g = lane >> 2, t = lane & 3;
    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    const float* pa[MT][2];
    const float* pb[NT];
    const int K = tk.k0 + tk.k1;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int r = tk.n0 + 16 * i + 8 * hh + g;
            pa[i][hh] = S + (r < tk.N ? tk.a_off + r * TMS : tk.zero_off) + t;
        }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int f = tk.kf0 + 8 * j + g;
        const int off = f < tk.k0 ? tk.in_off + f * TMS : (f < K ? tk.in1_off + (f - tk.k0) * TMS : (f == K ? tk.one_off : tk.zero_off));
        pb[j] = S + off + t;
    }
#pragma unroll 2
    for (int s0 = 0; s0 < TM; s0 += 8) {
        uint32_t a[MT][4], al[X3 ? MT : 1][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const float v[4] = {pa[i][0][s0], pa[i][1][s0], pa[i][0][s0 + 4], pa[i][1][s0 + 4]};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (X3) {
                    a[i][e] = m_rna(v[e]);
                    al[i][e] = m_rna(v[e] - m_trunc_tf32(a[i][e]));
                } else {
                    a[i][e] = m_bits(v[e]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float v0 = pb[j][s0], v1 = pb[j][s0 + 4];
            uint32_t b0 = m_bits(v0), b1 = m_bits(v1), l0 = 0, l1 = 0;
            if (X3) {
                b0 = m_rna(v0); b1 = m_rna(v1);
                l0 = m_rna(v0 - m_trunc_tf32(b0)); l1 = m_rna(v1 - m_trunc_tf32(b1));
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                if (X3) {
                    m_mma(acc[i][j], al[i], b0, b1);
                    m_mma(acc[i][j], a[i], l0, l1);
                }
                m_mma(acc[i][j], a[i], b0, b1);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int r = tk.n0 + 16 * i + 8 * hh + g;
            if (r < tk.N) {
                float* qr = partial + tk.out_off + r * tk.ld + tk.kf0 + 2 * t;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (tk.kf0 + 8 * j + 2 * t < tk.nstore) {   // nstore and ld are such that the pair stays inside the row
                        const float v0 = acc[i][j][2 * hh], v1 = acc[i][j][2 * hh + 1];
                        if (first) { qr[8 * j] = v0; qr[8 * j + 1] = v1; }
                        else m_red2(qr + 8 * j, v0, v1);
                    }
                }
            }
        }
    }
}

// ---- op record / coupling table fetch (kernel parameter when the program fits it, else global memory) ---------------------
HINT_DEV void m_ld_op(const MmaTables& T, const MmaParamProg& P, int idx, int (&r)[8]) {
    if (T.in_param) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = P.ops[idx].w[i];
        return;
    }
    const WOp* p = T.prog + idx;
#if defined(__CUDA_ARCH__)
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(p)), q1 = __ldg(reinterpret_cast<const int4*>(p) + 1);
    r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w; r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
#else
    for (int i = 0; i < 8; ++i) r[i] = p->w[i];
#endif
}
HINT_DEV Ep m_ld_ep(const MmaTables& T, const MmaParamProg& P, int e) {
    if (T.in_param) {
        Ep ep;
        ep.x_col = P.eps[e][0]; ep.s_col = P.eps[e][1]; ep.t_col = P.eps[e][2]; ep.pad = 0;
        return ep;
    }
    return T.eps[e];
}

// ---- tile I/O and coupling epilogues (thread -> sample mapping as in simt_phases.cuh, for kMmaThreads threads) --------
template <int TM>
HINT_DEV void m_load_tile(int tid, float* S, int col_base, const float* __restrict__ gsrc, long long row0, long long B, int width) {
    constexpr int TMS = TM + 4;
    if (width == 0) return;
    const long long base = row0 * width;
    const long long rows = (B - row0) < TM ? (B - row0) : TM;
    const int nvalid = (int)(rows * width);
    const int n = TM * width;
    for (int i = tid * 4; i < n; i += kMmaThreads * 4) {
        float v[4];
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            const float4 q = __ldg(reinterpret_cast<const float4*>(gsrc + base + i));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#else
            for (int e = 0; e < 4; ++e) v[e] = gsrc[base + i + e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (i + e < nvalid) ? gsrc[base + i + e] : 0.f;
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
HINT_DEV void m_store_tile(int tid, const float* S, int col_base, float* __restrict__ gdst, long long row0, long long B, int width) {
    constexpr int TMS = TM + 4;
    if (width == 0) return;
    const long long base = row0 * width;
    const long long rows = (B - row0) < TM ? (B - row0) : TM;
    const int nvalid = (int)(rows * width);
    for (int i = tid * 4; i < nvalid; i += kMmaThreads * 4) {
        float v[4];
        int m = i / width, j = i - m * width;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = (i + e < nvalid) ? S[(col_base + j) * TMS + m] : 0.f;
            if (++j == width) { j = 0; ++m; }
        }
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<float4*>(gdst + base + i) = make_float4(v[0], v[1], v[2], v[3]);
#else
            for (int e = 0; e < 4; ++e) gdst[base + i + e] = v[e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < nvalid) gdst[base + i + e] = v[e];
        }
    }
}

// hint.py:79-84.  forward: x_l <- e(s) x_l + t, J += alpha atan s;  inverse: x_l <- (x_l - t) / e(s), J -= alpha atan s
template <int TM>
HINT_DEV void m_coupling(int tid, float* S, const MmaTables& T, const MmaParamProg& P, int begin, int end, float alpha, int col_x,
                         float* JP, bool rev) {
    constexpr int TMS = TM + 4, NJG = kMmaThreads / TM;
    const int m = tid % TM;
    float jacc = 0.f;
    for (int e = begin + tid / TM; e < end; e += NJG) {
        const Ep ep = m_ld_ep(T, P, e);
        const float s = S[ep.s_col * TMS + m];
        const float t = S[ep.t_col * TMS + m];
        const float la = alpha * m_atan(s);
        float* xp = S + (col_x + ep.x_col) * TMS + m;
        if (!rev) {
            *xp = fmaf(m_exp(la), *xp, t);
            jacc += la;
        } else {
            *xp = (*xp - t) * m_exp(-la);
            jacc -= la;
        }
    }
    JP[tid] += jacc;
}

// backward coupling: x_l' = (z_l - t)/e ; dt = dz_l ; ds = (dz_l (z_l - t) + dJ) alpha/(1+s^2) ; dx_l' = dz_l e
template <int TM, bool ROUND>
HINT_DEV void m_coupling_bwd(int tid, float* S, const MmaTables& T, const MmaParamProg& P, int begin, int end, float alpha,
                             int col_x, int col_d, const float* DJ) {
    constexpr int TMS = TM + 4, NJG = kMmaThreads / TM;
    const int m = tid % TM;
    const float dj = DJ[m];
    for (int e = begin + tid / TM; e < end; e += NJG) {
        const Ep ep = m_ld_ep(T, P, e);
        float* sp = S + ep.s_col * TMS + m;
        float* tp = S + ep.t_col * TMS + m;
        float* xp = S + (col_x + ep.x_col) * TMS + m;
        float* dp = S + (col_d + ep.x_col) * TMS + m;
        const float s = *sp, t = *tp;
        const float la = alpha * m_atan(s);
        const float zl = *xp, dzl = *dp;
        const float r = zl - t;
        *xp = r * m_exp(-la);
        *dp = dzl * m_exp(la);
        const float ds = (dzl * r + dj) * (alpha * m_rcp(fmaf(s, s, 1.f)));
        *tp = ROUND ? m_round_tf32(dzl) : dzl;    // ds, dt are only ever GEMM operands (G3, dW3)
        *sp = ROUND ? m_round_tf32(ds) : ds;
    }
}

// T.begin[warp] without dynamically indexing the kernel-parameter array (which would force a local-memory copy)
HINT_DEV int m_begin(const MmaTables& T, int warp) {
    int b = T.begin[0];
#pragma unroll
    for (int w = 1; w < kMmaWarps; ++w) b = (warp == w) ? T.begin[w] : b;
    return b;
}

// ---- one tile through the tree: the warp interprets its op stream (plan_mma.h), next record always prefetched ------------
#define HINT_R8(r) r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]
template <int TM, bool X3>
HINT_DEV void m_dw_dispatch(const int (&r)[8], const float* S, float* __restrict__ partial, bool first, int lane) {
    const int mt = m_dw_mt(r), nt = m_dw_nt(r);
    if (mt == 2) {
        if (nt == 4) m_dw<TM, X3, 2, 4>(HINT_R8(r), S, partial, first, lane);
        else if (nt == 3) m_dw<TM, X3, 2, 3>(HINT_R8(r), S, partial, first, lane);
        else if (nt == 2) m_dw<TM, X3, 2, 2>(HINT_R8(r), S, partial, first, lane);
        else m_dw<TM, X3, 2, 1>(HINT_R8(r), S, partial, first, lane);
    } else {
        if (nt == 4) m_dw<TM, X3, 1, 4>(HINT_R8(r), S, partial, first, lane);
        else if (nt == 3) m_dw<TM, X3, 1, 3>(HINT_R8(r), S, partial, first, lane);
        else if (nt == 2) m_dw<TM, X3, 1, 2>(HINT_R8(r), S, partial, first, lane);
        else m_dw<TM, X3, 1, 1>(HINT_R8(r), S, partial, first, lane);
    }
}

// The interpreter.  BWD selects the backward coupling and enables the weight-gradient ops; `rev` the inverse coupling.
// `cur` holds the warp's first record (loaded by the caller before the tile load).
template <int TM, bool X3, bool BWD>
HINT_DEV void m_run_program(const MmaTables& T, const MmaParamProg& P, float* S, const float* __restrict__ W,
                            const float* __restrict__ Wlo, float* __restrict__ partial, bool first, int tid, int rev,
                            int (&cur)[8], long long* dbg) {
    const int warp = tid >> 5, lane = tid & 31;
    int pc = m_begin(T, warp);
    float* RAW = S + T.raw_off;
    BRing<X3, RingDepth<TM>::value> R;
    bool ring_valid = false;   // R holds the first fragments of the next OP_GEMM
    for (;;) {
        const int type = m_op_type(cur);
        if (type == OP_END) break;
        int nxt[8];
        m_ld_op(T, P, pc + 1, nxt);
        if (!T.in_param && lane == 0) m_prefetch_l1(T.prog + pc + 8);     // keep the op stream a few records ahead in L1
        const bool sync = m_op_flags(cur) & MT_SYNC;
        if (type == OP_GEMM) {
            const int nt = m_gemm_nt(cur);
            if (nt == 3) m_gemm_task<TM, X3, 3>(cur, nxt, sync, ring_valid, S, W, Wlo, lane, R, dbg, tid);
            else if (nt == 2) m_gemm_task<TM, X3, 2>(cur, nxt, sync, ring_valid, S, W, Wlo, lane, R, dbg, tid);
            else m_gemm_task<TM, X3, 1>(cur, nxt, sync, ring_valid, S, W, Wlo, lane, R, dbg, tid);
        } else {
            if (!ring_valid && m_op_type(nxt) == OP_GEMM) {
                m_issue_b_any<X3, RingDepth<TM>::value>(nxt, W, Wlo, lane, R);
                ring_valid = true;
            }
            if (sync || type == OP_SYNC) { m_cta_sync(); m_dbg_stamp(dbg, tid, 0); }
            if (type == OP_DW) {
                if (BWD) m_dw_dispatch<TM, X3>(cur, S, partial, first, lane);
            } else if (type == OP_COUPLE) {
                if (BWD) m_coupling_bwd<TM, !X3>(tid, S, T, P, cur[0], cur[1], T.alpha, T.col_x, T.col_d, RAW);
                else m_coupling<TM>(tid, S, T, P, cur[0], cur[1], T.alpha, T.col_x, RAW, rev != 0);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        ++pc;
    }
}

template <int TM, bool X3>
HINT_DEV void m_fwd_tile(const MmaTables& T, const MmaParamProg& P, float* S, const float* __restrict__ x, const float* __restrict__ c,
                         const float* __restrict__ W, const float* __restrict__ Wlo, float* __restrict__ z,
                         float* __restrict__ logdet, long long B, int rev, long long row0, int tid, long long* dbg) {
    float* JP = S + T.raw_off;
    m_dbg_stamp(dbg, tid, 0);
    int cur[8];
    m_ld_op(T, P, m_begin(T, tid >> 5), cur);
    m_load_tile<TM>(tid, S, T.col_x, x, row0, B, T.d);
    m_load_tile<TM>(tid, S, T.col_x + T.d, c, row0, B, T.dc);
    JP[tid] = 0.f;
    m_cta_sync();
    m_run_program<TM, X3, false>(T, P, S, W, Wlo, nullptr, false, tid, rev, cur, dbg);
    m_store_tile<TM>(tid, S, T.col_x, z, row0, B, T.d);
    if (tid < TM && row0 + tid < B) {
        float j = 0.f;
        for (int q = 0; q < kMmaThreads / TM; ++q) j += JP[q * TM + tid];
        logdet[row0 + tid] = j;
    }
    m_cta_sync();
}

template <int TM, bool X3>
HINT_DEV void m_bwd_tile(const MmaTables& T, const MmaParamProg& P, float* S, const float* __restrict__ z, const float* __restrict__ c,
                         const float* __restrict__ W, const float* __restrict__ Wlo, const float* __restrict__ dz,
                         const float* __restrict__ dlogdet, float* __restrict__ x_rec, float* __restrict__ dx,
                         float* __restrict__ dc, float* __restrict__ partial, bool first, long long B, long long row0, int tid,
                         long long* dbg) {
    constexpr int TMS = TM + 4;
    float* DJ = S + T.raw_off;
    m_dbg_stamp(dbg, tid, 0);
    int cur[8];
    m_ld_op(T, P, m_begin(T, tid >> 5), cur);
    m_load_tile<TM>(tid, S, T.col_x, z, row0, B, T.d);
    m_load_tile<TM>(tid, S, T.col_x + T.d, c, row0, B, T.dc);
    m_load_tile<TM>(tid, S, T.col_d, dz, row0, B, T.d);
    for (int i = tid; i < T.dc * TM; i += kMmaThreads) S[(T.col_d + T.d + i / TM) * TMS + i % TM] = 0.f;
    if (tid < TM) DJ[tid] = (row0 + tid < B) ? dlogdet[row0 + tid] : 0.f;
    m_cta_sync();
    m_run_program<TM, X3, true>(T, P, S, W, Wlo, partial, first, tid, 0, cur, dbg);
    if (x_rec) m_store_tile<TM>(tid, S, T.col_x, x_rec, row0, B, T.d);
    m_store_tile<TM>(tid, S, T.col_d, dx, row0, B, T.d);
    if (dc) m_store_tile<TM>(tid, S, T.col_d + T.d, dc, row0, B, T.dc);
    m_cta_sync();
}

template <int TM>
HINT_DEV void m_init_consts(const MmaTables& T, float* S, int tid) {
    constexpr int TMS = TM + 4;
    for (int i = tid; i < TMS; i += kMmaThreads) {
        S[T.col_one * TMS + i] = 1.f;
        S[T.col_zero * TMS + i] = 0.f;
    }
}

// whole-CTA bodies (persistent over tiles); `bid`/`nblocks` are blockIdx.x / gridDim.x
template <int TM, bool X3>
HINT_DEV void m_fwd_body(const MmaTables& T, const MmaParamProg& P, float* S, const float* __restrict__ x, const float* __restrict__ c,
                         const float* __restrict__ W, const float* __restrict__ Wlo, float* __restrict__ z,
                         float* __restrict__ logdet, long long B, int rev, int tid, int bid, int nblocks) {
    m_init_consts<TM>(T, S, tid);
    W += (bid % T.wcopies) * T.wstride;
    if (X3) Wlo += (bid % T.wcopies) * T.wstride;
    const long long ntiles = (B + TM - 1) / TM;
    for (long long tile = bid; tile < ntiles; tile += nblocks)
        m_fwd_tile<TM, X3>(T, P, S, x, c, W, Wlo, z, logdet, B, rev, tile * TM, tid, (bid == 0 && tile == nblocks) ? T.dbg : nullptr);
}

template <int TM, bool X3>
HINT_DEV void m_bwd_body(const MmaTables& T, const MmaParamProg& P, float* S, const float* __restrict__ z, const float* __restrict__ c,
                         const float* __restrict__ W, const float* __restrict__ Wlo, const float* __restrict__ dz,
                         const float* __restrict__ dlogdet, float* __restrict__ x_rec, float* __restrict__ dx,
                         float* __restrict__ dc, float* __restrict__ partials, long long n_partial, long long B, int tid,
                         int bid, int nblocks) {
    m_init_consts<TM>(T, S, tid);
    W += (bid % T.wcopies) * T.wstride;
    if (X3) Wlo += (bid % T.wcopies) * T.wstride;
    float* partial = partials + (long long)bid * n_partial;
    const long long ntiles = (B + TM - 1) / TM;
    bool first = true;
    for (long long tile = bid; tile < ntiles; tile += nblocks) {
        m_bwd_tile<TM, X3>(T, P, S, z, c, W, Wlo, dz, dlogdet, x_rec, dx, dc, partial, first, B, tile * TM, tid,
                           (bid == 0 && tile == nblocks) ? T.dbg : nullptr);
        first = false;
    }
}

#if defined(__CUDACC__)
template <int TM, bool X3>
__global__ void __launch_bounds__(kMmaThreads, HINT_MMA_MINB)
hint_fwd_mma_kernel(const __grid_constant__ MmaTables T, const __grid_constant__ MmaParamProg P, const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ W,
                    const float* __restrict__ Wlo, float* __restrict__ z, float* __restrict__ logdet, long long B, int rev) {
    extern __shared__ float4 m_smem4[];
    m_fwd_body<TM, X3>(T, P, reinterpret_cast<float*>(m_smem4), x, c, W, Wlo, z, logdet, B, rev, threadIdx.x, blockIdx.x, gridDim.x);
}

template <int TM, bool X3>
__global__ void __launch_bounds__(kMmaThreads, HINT_MMA_MINB)
hint_bwd_mma_kernel(const __grid_constant__ MmaTables T, const __grid_constant__ MmaParamProg P, const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ W,
                    const float* __restrict__ Wlo, const float* __restrict__ dz, const float* __restrict__ dlogdet,
                    float* __restrict__ x_rec, float* __restrict__ dx, float* __restrict__ dc, float* __restrict__ partials,
                    long long n_partial, long long B) {
    extern __shared__ float4 m_smem4[];
    m_bwd_body<TM, X3>(T, P, reinterpret_cast<float*>(m_smem4), z, c, W, Wlo, dz, dlogdet, x_rec, dx, dc, partials, n_partial, B,
                       threadIdx.x, blockIdx.x, gridDim.x);
}

static __global__ void hint_pack_mma_kernel(const int* __restrict__ src, const float* __restrict__ params, float* __restrict__ hi,
                                     float* __restrict__ lo, long long n, int copies, long long stride) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float h, l;
        m_pack_elem(src[i], params, h, l);
        for (int c = 0; c < copies; ++c) {
            hi[i + c * stride] = h;
            if (lo) lo[i + c * stride] = l;
        }
    }
}
#endif  // __CUDACC__

}  // namespace hint
