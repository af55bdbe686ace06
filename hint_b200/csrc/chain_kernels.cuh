// Register-chained warp-MMA kernels (HINT_MODE_TF32 when the block fits plan_chain.h's shapes): forward / inverse
// transport + log-det of one HINT coupling block (hint.py:62-101), subnets (hint.py:10-13) on mma.sync.m16n8k8 tf32 with
// fp32 accumulation.  Design and layouts: plan_chain.h.
//
// A warp owns RW = 16*MT samples.  MMA row <-> sample mapping (samples are independent, so any mapping works as long as
// the A fragments, the C fragments and the coupling agree): lane group g owns samples R*g .. R*g+R-1 (R = 2*MT), sample
// R*g + 2*i + hh is row g + 8*hh of m-tile i.  One column of the lane's rows is therefore ONE 64/128-bit shared-memory
// access, bank-conflict free at pitch RW+4 for the column pairs (2t, 2t+1) the fragments use.
//
// Same primitives as mma_kernels.cuh, so the same code runs under the fiber emulation (tests/emul/emul_chain.cpp).
#pragma once
#include "mma_kernels.cuh"
#include "plan_chain.h"

#if !defined(__CUDACC__)
namespace hint { namespace emu { void warp_sync(); void ldsm4(const float* rowp, uint32_t (&r)[4]); float shfl_xor(float v, int mask); } }
#endif

namespace hint {

// Developer experiments (timing only, some give WRONG results) are compiled in only with -DHINT_B200_DEV; in the product build
// HINT_EXP() is the constant 0, so the branches vanish and no environment variable can reach them.
#ifdef HINT_B200_DEV
#define HINT_EXP(e, bit) ((e) & (bit))
#else
#define HINT_EXP(e, bit) 0
#endif

// L2 residency hints.  The backward streams z / dz in and dx out once (evict-first) while the per-CTA partial-gradient buffers
// are re-accumulated after every tile (evict-last): without the hints the streamed tiles push the 85 MB of partials out of the
// 126 MB L2 and every flush goes to DRAM (round 1: 5.07 GB of DRAM traffic per launch for 0.55 GB of algorithmic bytes).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long c_policy_evict_first() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long c_policy_evict_last() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif

struct ChainTables {
    int n_nodes, d, dc;
    float alpha;
    int n_fwd_packed;   // floats of the forward operand region (staged in shared memory by the WS kernels)
    int exp;            // developer experiments (timing only, WRONG results): 1 no partial flush, 2 no dW GEMMs, 4 all operand
                        // loads hit the first 4 KB of the packed buffer (HINT_B200_CHAIN_EXP)
    float nll_scale;    // backward only, used when dz == NULL: the upstream gradient is that of the NLL loss of
                        // train_unconditional.py:128-132, generated in the tile load: dz = nll_scale * z; dlogdet == NULL: dlogdet = -nll_scale
};

HINT_DEV float c_shfl_xor(float v, int mask) {
#if defined(__CUDA_ARCH__)
    return __shfl_xor_sync(0xffffffffu, v, mask);
#elif !defined(__CUDACC__)
    return emu::shfl_xor(v, mask);
#else
    return v;
#endif
}

HINT_DEV void c_syncwarp() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#elif !defined(__CUDACC__)
    emu::warp_sync();
#endif
}

// weights: read-only, cached in L1 (every warp of the SM re-reads the same 100-200 KB of operands for every tile)
template <bool WS>
HINT_DEV void c_ldw2(const float* __restrict__ p, float& a, float& b) {
#if defined(__CUDA_ARCH__)
    if (WS) {   // operands staged in shared memory
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        a = v.x; b = v.y;
        return;
    }
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}

// the lane's R rows of one shared-memory column
template <int MT>
HINT_DEV void c_ld_rows(const float* p, float (&v)[2 * MT]) {
#if defined(__CUDA_ARCH__)
    if (MT == 2) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2 * MT - 2] = q.z; v[2 * MT - 1] = q.w;
    } else {
        const float2 q = *reinterpret_cast<const float2*>(p);
        v[0] = q.x; v[1] = q.y;
    }
#else
    for (int e = 0; e < 2 * MT; ++e) v[e] = p[e];
#endif
}
template <int MT>
HINT_DEV void c_st_rows(float* p, const float (&v)[2 * MT]) {
#if defined(__CUDA_ARCH__)
    if (MT == 2) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2 * MT - 2], v[2 * MT - 1]);
    else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
#else
    for (int e = 0; e < 2 * MT; ++e) p[e] = v[e];
#endif
}

// d = a*b + c with c in its own registers: the bias enters as the C operand of the first k-step, so no accumulator is ever
// initialised with moves
HINT_DEV void c_mma_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
#if defined(__CUDA_ARCH__)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
#else
    d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; d[3] = c[3];
    m_mma(d, a, b0, b1);
#endif
}
// d = a*b (zero accumulator: ptxas encodes the C operand as RZ)
HINT_DEV void c_mma_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if defined(__CUDA_ARCH__)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
#else
    d[0] = d[1] = d[2] = d[3] = 0.f;
    m_mma(d, a, b0, b1);
#endif
}
HINT_DEV uint32_t c_relu_rna(float v) { return m_bits(fmaxf(v, 0.f)) + 0x1000u; }

// One dense layer in registers: out[MT][NT_OUT] C fragments = bias + A[KS k-steps] * B.  B fragments (KS x NT_OUT) at Wl,
// bias (natural order) at bl.
template <bool WS, int MT, int KS, int NT_OUT>
HINT_DEV void c_layer(const uint32_t (&a)[KS][MT][4], const float* __restrict__ Wl, const float* __restrict__ bl, int lane,
                      float (&acc)[NT_OUT][MT][4]) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT_OUT; ++j) {
        float w0, w1;
        c_ldw2<WS>(Wl + j * 64 + 2 * lane, w0, w1);
        if (bl != nullptr) {
            float bq[4];
            c_ldw2<WS>(bl + 8 * j + 2 * t, bq[0], bq[1]);
            bq[2] = bq[0]; bq[3] = bq[1];
#pragma unroll
            for (int i = 0; i < MT; ++i) c_mma_c(acc[j][i], a[0][i], m_bits(w0), m_bits(w1), bq);
        } else {
#pragma unroll
            for (int i = 0; i < MT; ++i) c_mma_z(acc[j][i], a[0][i], m_bits(w0), m_bits(w1));
        }
    }
#pragma unroll
    for (int ks = 1; ks < KS; ++ks) {
#pragma unroll
        for (int j = 0; j < NT_OUT; ++j) {
            float w0, w1;
            c_ldw2<WS>(Wl + (ks * NT_OUT + j) * 64 + 2 * lane, w0, w1);
#pragma unroll
            for (int i = 0; i < MT; ++i) m_mma(acc[j][i], a[ks][i], m_bits(w0), m_bits(w1));
        }
    }
}

// block-diagonal layer (super node): n-tile j only sees k-step j; the operand holds the NT diagonal fragments
template <bool WS, int MT, int NT>
HINT_DEV void c_layer_bd(const uint32_t (&a)[NT][MT][4], const float* __restrict__ Wl, const float* __restrict__ bl, int lane,
                         float (&acc)[NT][MT][4]) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float w0, w1;
        c_ldw2<WS>(Wl + j * 64 + 2 * lane, w0, w1);
        if (bl != nullptr) {
            float bq[4];
            c_ldw2<WS>(bl + 8 * j + 2 * t, bq[0], bq[1]);
            bq[2] = bq[0]; bq[3] = bq[1];
#pragma unroll
            for (int i = 0; i < MT; ++i) c_mma_c(acc[j][i], a[j][i], m_bits(w0), m_bits(w1), bq);
        } else {
#pragma unroll
            for (int i = 0; i < MT; ++i) c_mma_z(acc[j][i], a[j][i], m_bits(w0), m_bits(w1));
        }
    }
}
template <bool WS, int MT, int NT, int BD>
HINT_DEV void c_layer_hh(const uint32_t (&a)[NT][MT][4], const float* __restrict__ Wl, const float* __restrict__ bl, int lane,
                         float (&acc)[NT][MT][4]) {
    if (BD) c_layer_bd<WS, MT, NT>(a, Wl, bl, lane, acc);
    else c_layer<WS, MT, NT, NT>(a, Wl, bl, lane, acc);
}

// C fragments -> A fragments of the next layer (k-slot permutation: a = {c0, c2, c1, c3}), ReLU and tf32 rounding in place
template <int MT, int NT>
HINT_DEV void c_relu_to_a(const float (&acc)[NT][MT][4], uint32_t (&a)[NT][MT][4]) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            a[j][i][0] = c_relu_rna(acc[j][i][0]); a[j][i][1] = c_relu_rna(acc[j][i][2]);
            a[j][i][2] = c_relu_rna(acc[j][i][1]); a[j][i][3] = c_relu_rna(acc[j][i][3]);
        }
}

// layer-1 A fragments from the warp's x tile: input feature f reads tile column in_col[f] (x column, or d + j for
// condition column j); -1 = no such feature (zero; the packed W1 rows are zero there too)
template <int MT, int KS1>
HINT_DEV void c_load_input(const float* XT, const short* in_col, int lane, uint32_t (&a)[KS1][MT][4]) {
    constexpr int PW = 16 * MT + 4, R = 2 * MT;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = in_col[8 * ks + 2 * t + c];
            float v[R];
#pragma unroll
            for (int e = 0; e < R; ++e) v[e] = 0.f;
            if (col >= 0) c_ld_rows<MT>(XT + col * PW + R * g, v);
#pragma unroll
            for (int i = 0; i < MT; ++i) { a[ks][i][2 * c] = m_rna(v[2 * i]); a[ks][i][2 * c + 1] = m_rna(v[2 * i + 1]); }
        }
}

// s / t subnet of one (super) node (hint.py:10-13,77) entirely in registers
template <bool WS, int MT, int KS1, int NH, int NO, int BD>
HINT_DEV void c_subnet(const uint32_t (&a1)[KS1][MT][4], const float* __restrict__ Wnet, int lane, float (&out)[NO][MT][4]) {
    using O = ChainOff<KS1, NH, NO, BD>;
    uint32_t h[NH][MT][4];
    {
        float acc[NH][MT][4];
        c_layer<WS, MT, KS1, NH>(a1, Wnet + O::w1, Wnet + O::b1, lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    {
        float acc[NH][MT][4];
        c_layer_hh<WS, MT, NH, BD>(h, Wnet + O::w2, Wnet + O::b2, lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    c_layer<WS, MT, NH, NO>(h, Wnet + O::w3, Wnet + O::b3, lane, out);
}

// one (super) node, forward (hint.py:79-81) or inverse (hint.py:82-84) coupling; JP = the lane's private log-det partials
template <bool WS, int MT, int KS1, int NH, int NO, int BD, bool REV>
HINT_DEV void c_node_fwd(const ChainNode* nd, float alpha, const float* __restrict__ Wn, float* XT, float (&jl)[2 * MT], int lane) {
    constexpr int PW = 16 * MT + 4, R = 2 * MT;
    using O = ChainOff<KS1, NH, NO, BD>;
    const int g = lane >> 2, t = lane & 3;
    uint32_t a1[KS1][MT][4];
    c_load_input<MT, KS1>(XT, nd->in_col, lane, a1);
    float s[NO][MT][4], tt[NO][MT][4];
    c_subnet<WS, MT, KS1, NH, NO, BD>(a1, Wn, lane, s);
    c_subnet<WS, MT, KS1, NH, NO, BD>(a1, Wn + O::net, lane, tt);
#pragma unroll
    for (int j = 0; j < NO; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int xcol = nd->out_col[8 * j + 2 * t + c];
            if (xcol >= 0) {
                float* xp = XT + xcol * PW + R * g;
                float xv[R];
                c_ld_rows<MT>(xp, xv);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float la = alpha * m_atan(s[j][i][2 * hh + c]);
                        const float tv = tt[j][i][2 * hh + c];
                        float& x = xv[2 * i + hh];
                        if (!REV) { x = fmaf(m_exp(la), x, tv); jl[2 * i + hh] += la; }
                        else { x = (x - tv) * m_exp(-la); jl[2 * i + hh] -= la; }
                    }
                c_st_rows<MT>(xp, xv);
            }
        }
}

#define HINT_CHAIN_SHAPES(X) X(0, 1, 1, 1, 0) X(1, 1, 2, 1, 1) X(2, 1, 4, 1, 1) X(3, 1, 2, 1, 0) X(4, 1, 3, 1, 0) \
                             X(5, 1, 5, 1, 0) X(6, 2, 5, 2, 0) X(7, 2, 9, 2, 0) X(8, 3, 9, 3, 0)

template <bool WS, int MT, bool REV>
HINT_DEV void c_node_fwd_dispatch(const ChainNode* nd, float alpha, const float* __restrict__ W, float* XT, float (&JP)[2 * MT], int lane) {
    const float* Wn = W + nd->w_off;
#define HINT_CHAIN_CASE(ID, A, B, C, D) \
    case ID: c_node_fwd<WS, MT, A, B, C, D, REV>(nd, alpha, Wn, XT, JP, lane); break;
    switch (nd->shape) {
        HINT_CHAIN_SHAPES(HINT_CHAIN_CASE)
        default: break;
    }
#undef HINT_CHAIN_CASE
}

// warp-private tile I/O: global rows [row0, row0+rows) x width floats (row-major, contiguous) <-> columns of the tile
template <int MT>
HINT_DEV void c_load_tile(float* XT, int col_base, const float* __restrict__ gsrc, long long row0, int rows, int width, int lane) {
    constexpr int PW = 16 * MT + 4, RW = 16 * MT;
    if (width == 0) return;
    const float* src = gsrc + row0 * width;
    const int nvalid = rows * width, n = RW * width;
    int m0 = (lane * 4) / width, j0 = lane * 4 - m0 * width;      // (row, column) of element i, advanced without divisions
    const int dm = 128 / width, dj = 128 - dm * width;
    for (int i = lane * 4; i < n; i += 128) {
        float v[4];
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            // streamed once: keep the tile out of L1, which holds the weights
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(src + i), "l"(c_policy_evict_first()));
#else
            for (int e = 0; e < 4; ++e) v[e] = src[i + e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (i + e < nvalid) ? src[i + e] : 0.f;
        }
        int m = m0, j = j0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            XT[(col_base + j) * PW + m] = v[e];
            if (++j == width) { j = 0; ++m; }
        }
        m0 += dm; j0 += dj;
        if (j0 >= width) { j0 -= width; ++m0; }
    }
}
template <int MT>
HINT_DEV void c_store_tile(const float* XT, int col_base, float* __restrict__ gdst, long long row0, int rows, int width, int lane) {
    constexpr int PW = 16 * MT + 4;
    if (width == 0) return;
    float* dst = gdst + row0 * width;
    const int nvalid = rows * width;
    int m0 = (lane * 4) / width, j0 = lane * 4 - m0 * width;
    const int dm = 128 / width, dj = 128 - dm * width;
    for (int i = lane * 4; i < nvalid; i += 128) {
        float v[4];
        int m = m0, j = j0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = XT[(col_base + j) * PW + m];
            if (++j == width) { j = 0; ++m; }
        }
        m0 += dm; j0 += dj;
        if (j0 >= width) { j0 -= width; ++m0; }
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                         :: "l"(dst + i), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "l"(c_policy_evict_first()) : "memory");
#else
            for (int e = 0; e < 4; ++e) dst[i + e] = v[e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < nvalid) dst[i + e] = v[e];
        }
    }
}

// floats of one warp's private shared-memory region (x + condition columns)
template <int MT>
HINT_HD constexpr int chain_fwd_warp_floats(int d, int dc) { return (d + dc) * (16 * MT + 4); }

template <int MT, int NW, bool REV, bool WS>
HINT_DEV void c_fwd_body(const ChainTables& T, const ChainNode* nodes, float* S, const float* __restrict__ x, const float* __restrict__ c,
                         const float* __restrict__ W, float* __restrict__ z, float* __restrict__ logdet, long long B, int tid,
                         int bid, int nblocks) {
    constexpr int RW = 16 * MT;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int R = 2 * MT;
    float* XT = S + warp * chain_fwd_warp_floats<MT>(T.d, T.dc);
    const long long ntiles = (B + RW - 1) / RW;
    // The warps are independent, but they all run the same (large, fully unrolled) instruction stream: a CTA barrier per
    // tile HINT_EXP(T.exp, 8) or per node HINT_EXP(T.exp, 16) keeps them in step so they share the instruction cache.
    for (long long base = (long long)bid * NW; base < ntiles; base += (long long)nblocks * NW) {
        const long long tile = base + warp;
        const bool live = tile < ntiles;
        const long long row0 = tile * RW;
        const int rows = !live ? 0 : (int)((B - row0) < RW ? (B - row0) : RW);
        if (live) {
            c_load_tile<MT>(XT, 0, x, row0, rows, T.d, lane);
            c_load_tile<MT>(XT, T.d, c, row0, rows, T.dc, lane);
            c_syncwarp();
        }
        float JP[R];      // the lane's log-det partials of its R rows (summed over the 4 lanes of a row group at the end)
#pragma unroll
        for (int e = 0; e < R; ++e) JP[e] = 0.f;
        for (int q = 0; q < T.n_nodes; ++q) {
            if (live) {
                c_node_fwd_dispatch<WS, MT, REV>(nodes + (REV ? T.n_nodes - 1 - q : q), T.alpha, W, XT, JP, lane);
                c_syncwarp();
            }
            if (HINT_EXP(T.exp, 16)) m_cta_sync();
        }
        if (live) {
            c_store_tile<MT>(XT, 0, z, row0, rows, T.d, lane);
#pragma unroll
            for (int e = 0; e < R; ++e) {
                JP[e] += c_shfl_xor(JP[e], 1);
                JP[e] += c_shfl_xor(JP[e], 2);
                const int row = R * (lane >> 2) + e;
                if ((lane & 3) == 0 && row < rows) logdet[row0 + row] = JP[e];
            }
            c_syncwarp();
        }
        if (HINT_EXP(T.exp, 8)) m_cta_sync();
    }
}

// shared memory of the forward kernels: [node table | forward operands (WS only) | NW warp tiles]
HINT_HD constexpr int chain_node_floats(int n_nodes) { return n_nodes * 32; }
template <int MT>
HINT_HD constexpr size_t chain_fwd_smem_bytes(int n_nodes, int d, int dc, int nw, long long n_fwd_packed, bool ws) {
    return 4 * ((size_t)chain_node_floats(n_nodes) + (ws ? (size_t)n_fwd_packed : 0) + (size_t)nw * chain_fwd_warp_floats<MT>(d, dc));
}

#if defined(__CUDACC__)
// WS: the forward operands are copied into shared memory once per CTA and every B fragment is an LDS (no L1 misses on the
// dependent chain of a node); otherwise they are read from global memory through L1.
template <int MT, int NW, bool REV, bool WS>
__global__ void __launch_bounds__(32 * NW, 1)
hint_fwd_chain_kernel(const __grid_constant__ ChainTables T, const __grid_constant__ ChainParam P, const float* __restrict__ x,
                      const float* __restrict__ c, const float* __restrict__ W, float* __restrict__ z, float* __restrict__ logdet,
                      long long B) {
    extern __shared__ float4 c_smem4[];
    float* S = reinterpret_cast<float*>(c_smem4);
    int* nodes = reinterpret_cast<int*>(S);
    for (int i = threadIdx.x; i < T.n_nodes * 32; i += 32 * NW) nodes[i] = reinterpret_cast<const int*>(P.nodes)[i];
    float* Ws = S + chain_node_floats(T.n_nodes);
    if (WS) {
        const float4* src = reinterpret_cast<const float4*>(W);
        float4* dst = reinterpret_cast<float4*>(Ws);
        for (int i = threadIdx.x; i < T.n_fwd_packed / 4; i += 32 * NW) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    float* tiles = Ws + (WS ? T.n_fwd_packed : 0);
    if (WS) c_fwd_body<MT, NW, REV, true>(T, reinterpret_cast<const ChainNode*>(nodes), tiles, x, c, Ws, z, logdet, B, threadIdx.x, blockIdx.x, gridDim.x);
    else c_fwd_body<MT, NW, REV, false>(T, reinterpret_cast<const ChainNode*>(nodes), tiles, x, c, W, z, logdet, B, threadIdx.x, blockIdx.x, gridDim.x);
}
#endif


// =====================================================================================================================
// Backward (memory-free; the autograd tape of hint.py:62-101): input is the block OUTPUT z.  Nodes are walked
// parent-first (the inverse order, hint.py:85-88).  A CTA owns a tile of TM = 16*MT*NW samples; the tile state lives in
// CTA-wide shared-memory columns [column][sample] with an XOR swizzle of the 16-byte chunks (c_swz) so that BOTH access
// patterns are bank-conflict free: a warp's private rows of the column pairs (2t, 2t+1) - the C/A fragment pattern of the
// register-chained subnets - and 4 consecutive samples of 8 consecutive columns - the operand pattern of the weight-gradient
// GEMMs, which contract over the tile's samples.  Per node and net:
//   private (each warp its own 16*MT samples, everything in registers): recompute the subnet, coupling backward
//     x_l' = (z_l - t)/e, dx_l' = dz_l e, ds = (dz_l (z_l - t) + dJ) alpha/(1+s^2), dt = dz_l, then the dgrad chain
//     dh2 = (dout W3) [h2>0], dh1 = (dh2 W2) [h1>0], dx_u' += dh1 W1; h1, h2, dh1, dh2, dout are also written (tf32) to the
//     CTA-wide buffers HB1, HB2, GB1, GB2, GO;
//   barrier; cooperative: dW3T += [h2|1]^T dout, dW2T += [h1|1]^T dh2, dW1T += [a|1]^T dh1 as m16n8k8 MMAs with K = the
//     TM samples of the tile, output tiles dealt round-robin to the warps, flushed to the CTA's private partial-gradient
//     buffer (store on the first tile, red.global.add afterwards; reduced in fixed order by hint_reduce_unpack_kernel);
//   barrier.
HINT_DEV int c_swz(int col) { return col & 7; }

// two matrices (lanes 0-15 supply the row addresses): one B fragment
HINT_DEV void c_ldsm2(const float* rowp, uint32_t (&r)[4]) {
#if defined(__CUDA_ARCH__)
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(rowp)) : "memory");
#elif !defined(__CUDACC__)
    emu::ldsm4(rowp, r);
#endif
}

// developer aid HINT_EXP(T.exp, 32): thread 0 of CTA 0 records clock64 at the phase boundaries of the backward sweep
#if defined(__CUDACC__)
#ifdef HINT_B200_DEV
__device__ long long g_chain_dbg[2048];
#endif
#endif
HINT_DEV void c_stamp(int exp, int warp, int lane) {
#if defined(__CUDA_ARCH__) && defined(HINT_B200_DEV)
    if (HINT_EXP(exp, 32) && warp == 0 && lane == 0 && blockIdx.x == 0) {
        const long long n = g_chain_dbg[0];
        if (n < 2040) { g_chain_dbg[1 + n] = clock64(); g_chain_dbg[0] = n + 1; }
    }
#endif
}

// ldmatrix of four 8x4 tf32 matrices (= 8x8 b16): lane l supplies the address of row l%8 (16 bytes) of matrix l/8 and
// receives element (row l/4, column l%4) of every matrix - exactly the m16n8k8 A fragment (matrices: rows 0-7 / 8-15 x
// k-slots 0-3 / 4-7) or two B fragments, in consecutive registers
HINT_DEV void c_ldsm4(const float* rowp, uint32_t (&r)[4]) {
#if defined(__CUDA_ARCH__)
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(rowp)));
#elif !defined(__CUDACC__)
    emu::ldsm4(rowp, r);
#endif
}

// float index of the lane's R = 2*MT rows (samples 16*MT*warp + R*g ..) of column `col`
template <int MT, int NW>
HINT_DEV int c_rows(int col, int warp, int g) {
    constexpr int TM = 16 * MT * NW;
    if (MT == 2) return col * TM + (((8 * warp + g) ^ c_swz(col)) << 2);
    return col * TM + (((4 * warp + (g >> 1)) ^ c_swz(col)) << 2) + 2 * (g & 1);
}
// float index of sample m of column col
template <int TM>
HINT_DEV int c_elem(int col, int m) { return col * TM + ((((m >> 2) ^ c_swz(col))) << 2) + (m & 3); }

struct ChainBwdSmem {      // float offsets of the CTA's shared-memory regions
    int xt, dz, hb1, hb2, gb1, gb2, go, dj, cst, nodes, total;
};
template <int MT, int NW>
HINT_HD constexpr ChainBwdSmem chain_bwd_smem(int d, int dc, int max_nh, int max_no, int n_nodes) {
    constexpr int TM = 16 * MT * NW;
    ChainBwdSmem s{};
    const int xc = (d + dc + 7) & ~7;     // buffers start at multiples of 8 columns (swizzle phase of the fragment columns)
    s.xt = 0;
    s.dz = s.xt + xc * TM;
    s.hb1 = s.dz + xc * TM;
    s.hb2 = s.hb1 + 8 * max_nh * TM;
    s.gb1 = s.hb2 + 8 * max_nh * TM;
    s.gb2 = s.gb1 + 8 * max_nh * TM;
    s.go = s.gb2 + 8 * max_nh * TM;
    s.dj = s.go + 8 * max_no * TM;
    s.cst = s.dj + TM;          // 4 x 1.0f, 4 x 0.0f: the bias row / padding rows of the weight-gradient operands
    s.nodes = s.cst + 8;
    s.total = s.nodes + 32 * n_nodes;
    return s;
}

template <int MT, int NW, int KS1>
HINT_DEV void c_load_input_sw(const float* XT, const short* in_col, int warp, int lane, uint32_t (&a)[KS1][MT][4]) {
    constexpr int R = 2 * MT;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = in_col[8 * ks + 2 * t + c];
            float v[R];
#pragma unroll
            for (int e = 0; e < R; ++e) v[e] = 0.f;
            if (col >= 0) c_ld_rows<MT>(XT + c_rows<MT, NW>(col, warp, g), v);
#pragma unroll
            for (int i = 0; i < MT; ++i) { a[ks][i][2 * c] = m_rna(v[2 * i]); a[ks][i][2 * c + 1] = m_rna(v[2 * i + 1]); }
        }
}

// A fragments (already ReLU'd and rounded) -> the warp's rows of a hidden buffer.  rp[c] = the lane's row address of buffer
// column 2t+c; n-tile j is 8*j*TM floats further.
template <int MT, int TM, int NT>
HINT_DEV void c_store_afrag(float* buf, const int (&rp)[2], const uint32_t (&a)[NT][MT][4]) {
    constexpr int R = 2 * MT;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v[R];
#pragma unroll
            for (int i = 0; i < MT; ++i) { v[2 * i] = m_float(a[j][i][2 * c]); v[2 * i + 1] = m_float(a[j][i][2 * c + 1]); }
            c_st_rows<MT>(buf + rp[c] + 8 * j * TM, v);
        }
}

// dgrad epilogue: G (C fragments) masked by the stored forward activation (> 0 <=> its rounded bits exceed the rounding
// increment), rounded to tf32 -> A fragments of the next dgrad layer
template <int MT, int TM, int NT>
HINT_DEV void c_mask_to_a(const float (&acc)[NT][MT][4], const float* hbuf, const int (&rp)[2], uint32_t (&a)[NT][MT][4]) {
    constexpr int R = 2 * MT;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float hv[R];
            c_ld_rows<MT>(hbuf + rp[c] + 8 * j * TM, hv);
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    a[j][i][2 * c + hh] = m_bits(hv[2 * i + hh]) > 0x1000u ? m_rna(acc[j][i][2 * hh + c]) : 0u;
        }
}

// weight-gradient GEMM of one layer: CT[in-feature rows | ones row][out cols] += sum over the tile's samples.  Both operands
// are fetched with ldmatrix from the swizzled [column][sample] buffers (a matrix row = 4 consecutive samples of one column).
//   XIN: the in-features are the node's subnet inputs (tile columns in_col[] of the x tile, exact fp32: the tensor core
//   truncates them), else columns of the buffer at `aoff`.
//   rot: rotates the unit -> warp assignment so that the small GEMMs of one phase land on different warps
template <int MT, int NW, int KSIN, int NTOUT, bool XIN>
HINT_DEV void c_dw_gemm(const float* S, int aoff, const short* in_col, int boff, int cst, float* __restrict__ part, bool first,
                        int warp, int lane, int rot, int exp) {
    if (HINT_EXP(exp, 2)) return;
    constexpr int TM = 16 * MT * NW;
    constexpr int MTC = (8 * KSIN + 1 + 15) / 16;
    constexpr int NC = NTOUT <= 5 ? NTOUT : 3;
    constexpr int NCH = NTOUT / NC;
    constexpr int NP = (NC + 1) / 2;
    static_assert(NCH * NC == NTOUT, "n-chunking must tile the layer");
    for (int u = (warp + NW - rot) % NW; u < MTC * NCH; u += NW) {
        const int i = u / NCH, ch = u - i * NCH;
        // A: matrix m = lane/8 -> rows 8*(m&1).. of the m-tile, chunk (m>>1) of the k-step
        int abase, amask = ~0;
        {
            const int r = 16 * i + 8 * ((lane >> 3) & 1) + (lane & 7), hi = lane >> 4;
            int col = -1;
            if (r < 8 * KSIN) col = XIN ? (int)in_col[r] : r;
            if (col >= 0) {
                const int sw = c_swz(col);
                abase = (XIN ? 0 : aoff) + col * TM + ((hi ^ (sw & 1)) << 2) + ((sw >> 1) << 3);
            } else {
                abase = cst + (r == 8 * KSIN ? 0 : 4);
                amask = 0;
            }
        }
        // B: pair p = n-tiles (2p, 2p+1) of the chunk; matrix m -> n-tile 2p + (m>>1), chunk (m&1)
        int bbase[NP], bmask[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int jj = 2 * p + (lane >> 4), hi = (lane >> 3) & 1;
            if (jj < NC) {
                const int col = 8 * (ch * NC + jj) + (lane & 7);
                const int sw = c_swz(col);
                bbase[p] = boff + col * TM + ((hi ^ (sw & 1)) << 2) + ((sw >> 1) << 3);
                bmask[p] = ~0;
            } else {
                bbase[p] = cst + 4;
                bmask[p] = 0;
            }
        }
        // operands of k-step ks+1 are fetched before the MMAs of k-step ks are issued (the asm statements keep program order,
        // so this IS the schedule): the ldmatrix latency overlaps the tensor pipe
        float acc[NC][4];
        uint32_t a[2][4], b[2][NP][4];
        c_ldsm4(S + abase, a[0]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (2 * p + 1 < NC) c_ldsm4(S + bbase[p], b[0][p]);
            else c_ldsm2(S + bbase[p], b[0][p]);       // odd last n-tile: half the fetch
        }
#pragma unroll
        for (int ks = 0; ks < TM / 8; ++ks) {
            const int cur = ks & 1, nxt = cur ^ 1;
            if (ks + 1 < TM / 8) {
                c_ldsm4(S + (abase ^ (((ks + 1) << 3) & amask)), a[nxt]);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    if (2 * p + 1 < NC) c_ldsm4(S + (bbase[p] ^ (((ks + 1) << 3) & bmask[p])), b[nxt][p]);
                    else c_ldsm2(S + (bbase[p] ^ (((ks + 1) << 3) & bmask[p])), b[nxt][p]);
                }
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (ks == 0) {
                    c_mma_z(acc[2 * p], a[cur], b[cur][p][0], b[cur][p][1]);
                    if (2 * p + 1 < NC) c_mma_z(acc[2 * p + 1], a[cur], b[cur][p][2], b[cur][p][3]);
                } else {
                    m_mma(acc[2 * p], a[cur], b[cur][p][0], b[cur][p][1]);
                    if (2 * p + 1 < NC) m_mma(acc[2 * p + 1], a[cur], b[cur][p][2], b[cur][p][3]);
                }
            }
        }
        // flush: C fragment (i, j) = 128 floats, the lane's 4 are contiguous
        if (HINT_EXP(exp, 1)) continue;
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
            float* q = part + ((i * NTOUT + ch * NC + jj) * 32 + lane) * 4;
#if defined(__CUDA_ARCH__)
            const unsigned long long keep = c_policy_evict_last();
            if (first) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                                    :: "l"(q), "f"(acc[jj][0]), "f"(acc[jj][1]), "f"(acc[jj][2]), "f"(acc[jj][3]), "l"(keep) : "memory");
            else asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                              :: "l"(q), "f"(acc[jj][0]), "f"(acc[jj][1]), "f"(acc[jj][2]), "f"(acc[jj][3]), "l"(keep) : "memory");
#else
            for (int e = 0; e < 4; ++e) { if (first) q[e] = acc[jj][e]; else q[e] += acc[jj][e]; }
#endif
        }
    }
}

// s / t subnet keeping h1 / h2 (tf32) in the CTA-wide buffers at hb1 / hb2; L3 = also produce the output
template <int MT, int TM, int KS1, int NH, int NO, int BD, bool L3>
HINT_DEV void c_subnet_keep(const uint32_t (&a1)[KS1][MT][4], const float* __restrict__ Wnet, int lane, float* hb1, float* hb2,
                            const int (&rp)[2], float (&out)[NO][MT][4]) {
    using O = ChainOff<KS1, NH, NO, BD>;
    uint32_t h[NH][MT][4];
    {
        float acc[NH][MT][4];
        c_layer<false, MT, KS1, NH>(a1, Wnet + O::w1, Wnet + O::b1, lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    c_store_afrag<MT, TM, NH>(hb1, rp, h);
    {
        float acc[NH][MT][4];
        c_layer_hh<false, MT, NH, BD>(h, Wnet + O::w2, Wnet + O::b2, lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    c_store_afrag<MT, TM, NH>(hb2, rp, h);
    if (L3) c_layer<false, MT, NH, NO>(h, Wnet + O::w3, Wnet + O::b3, lane, out);
}

// dgrad chain of one net: dh2 = (dout W3)[h2>0] -> gb2, dh1 = (dh2 W2)[h1>0] -> gb1, dx_upper / dc += dh1 W1
template <int MT, int NW, int KS1, int NH, int NO, int BD>
HINT_DEV void c_dgrad(const uint32_t (&dout)[NO][MT][4], const float* __restrict__ Wb, const short* in_col, float* DZ, const float* hb1,
                      const float* hb2, float* gb1, float* gb2, const int (&rp)[2], int warp, int lane) {
    constexpr int TM = 16 * MT * NW, R = 2 * MT;
    using O = ChainOff<KS1, NH, NO, BD>;
    const int g = lane >> 2, t = lane & 3;
    uint32_t gh[NH][MT][4];
    {
        float acc[NH][MT][4];
        c_layer<false, MT, NO, NH>(dout, Wb + O::w3t, nullptr, lane, acc);
        c_mask_to_a<MT, TM, NH>(acc, hb2, rp, gh);
    }
    c_store_afrag<MT, TM, NH>(gb2, rp, gh);
    {
        float acc[NH][MT][4];
        c_layer_hh<false, MT, NH, BD>(gh, Wb + O::w2t, nullptr, lane, acc);
        c_mask_to_a<MT, TM, NH>(acc, hb1, rp, gh);
    }
    c_store_afrag<MT, TM, NH>(gb1, rp, gh);
    float da[KS1][MT][4];
    c_layer<false, MT, NH, KS1>(gh, Wb + O::w1t, nullptr, lane, da);
#pragma unroll
    for (int j = 0; j < KS1; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = in_col[8 * j + 2 * t + c];
            if (col >= 0) {
                float* p = DZ + c_rows<MT, NW>(col, warp, g);
                float v[R];
                c_ld_rows<MT>(p, v);
#pragma unroll
                for (int i = 0; i < MT; ++i) { v[2 * i] += da[j][i][c]; v[2 * i + 1] += da[j][i][2 + c]; }
                c_st_rows<MT>(p, v);
            }
        }
}

// one (super) node of the backward sweep.  Nodes with NH <= 4 keep BOTH nets' activations in the buffers (net e at n-tile
// slots [e*NH, (e+1)*NH)): one private phase, one barrier pair; wider nodes run the nets one after the other and recompute
// the hidden activations of the t subnet.
template <int MT, int NW, int KS1, int NH, int NO, int BD>
HINT_DEV void c_node_bwd(const ChainNode* nd, float alpha, const float* __restrict__ Wn, const float* __restrict__ Wt,
                         float* S, const ChainBwdSmem& L, float* __restrict__ part, bool first, int warp, int lane, int exp) {
    constexpr int TM = 16 * MT * NW, R = 2 * MT;
    constexpr bool BOTH = NH <= 4;
    constexpr int HS = 8 * NH * TM, OS = 8 * NO * TM;      // floats between the two nets' slots
    using O = ChainOff<KS1, NH, NO, BD>;
    const int g = lane >> 2, t = lane & 3;
    float* XT = S + L.xt;
    float* DZ = S + L.dz;
    // the lane's row addresses of buffer columns 2t, 2t+1 (same in every buffer that starts at a multiple of 8 columns)
    int rp[2];
    rp[0] = c_rows<MT, NW>(2 * t, warp, g);
    rp[1] = c_rows<MT, NW>(2 * t + 1, warp, g);
    c_stamp(exp, warp, lane);
    uint32_t a1[KS1][MT][4];
    c_load_input_sw<MT, NW, KS1>(XT, nd->in_col, warp, lane, a1);
    uint32_t ds[NO][MT][4], dt[NO][MT][4];      // ds, dt as A fragments
    {
        float s[NO][MT][4], tt[NO][MT][4];
        if (BOTH) c_subnet_keep<MT, TM, KS1, NH, NO, BD, true>(a1, Wn + O::net, lane, S + L.hb1 + HS, S + L.hb2 + HS, rp, tt);
        else c_subnet<false, MT, KS1, NH, NO, BD>(a1, Wn + O::net, lane, tt);
        c_subnet_keep<MT, TM, KS1, NH, NO, BD, true>(a1, Wn, lane, S + L.hb1, S + L.hb2, rp, s);
        // coupling backward on the C fragments
        float dj[R];
        c_ld_rows<MT>(S + L.dj + 16 * MT * warp + R * g, dj);
#pragma unroll
        for (int j = 0; j < NO; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int xcol = nd->out_col[8 * j + 2 * t + c];
                float dsv[R], dtv[R];
#pragma unroll
                for (int e = 0; e < R; ++e) { dsv[e] = 0.f; dtv[e] = 0.f; }
                if (xcol >= 0) {
                    const int ra = c_rows<MT, NW>(xcol, warp, g);
                    float zl[R], dzl[R];
                    c_ld_rows<MT>(XT + ra, zl);
                    c_ld_rows<MT>(DZ + ra, dzl);
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int e = 2 * i + hh;
                            const float sv = s[j][i][2 * hh + c], tv = tt[j][i][2 * hh + c];
                            const float la = alpha * m_atan(sv);
                            const float r = zl[e] - tv;
                            zl[e] = r * m_exp(-la);
                            dsv[e] = (dzl[e] * r + dj[e]) * (alpha * m_rcp(fmaf(sv, sv, 1.f)));
                            dtv[e] = dzl[e];
                            dzl[e] = dzl[e] * m_exp(la);
                        }
                    c_st_rows<MT>(XT + ra, zl);
                    c_st_rows<MT>(DZ + ra, dzl);
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        ds[j][i][2 * c + hh] = m_rna(dsv[2 * i + hh]);
                        dt[j][i][2 * c + hh] = m_rna(dtv[2 * i + hh]);
                    }
            }
    }
    constexpr int rot2 = (O::mth * (NO <= 5 ? 1 : NO / 3)) % NW;
    constexpr int rot3 = (rot2 + O::mth * (NH <= 5 ? 1 : NH / 3)) % NW;
    constexpr int rotn = (rot3 + O::mt1 * (NH <= 5 ? 1 : NH / 3)) % NW;
    if (BOTH) {
        c_store_afrag<MT, TM, NO>(S + L.go, rp, ds);
        c_store_afrag<MT, TM, NO>(S + L.go + OS, rp, dt);
        c_dgrad<MT, NW, KS1, NH, NO, BD>(ds, Wt, nd->in_col, DZ, S + L.hb1, S + L.hb2, S + L.gb1, S + L.gb2, rp, warp, lane);
        c_dgrad<MT, NW, KS1, NH, NO, BD>(dt, Wt + O::tnet, nd->in_col, DZ, S + L.hb1 + HS, S + L.hb2 + HS, S + L.gb1 + HS, S + L.gb2 + HS,
                                         rp, warp, lane);
        c_stamp(exp, warp, lane);
        m_cta_sync();
        c_stamp(exp, warp, lane);
#pragma unroll 1
        for (int net = 0; net < 2; ++net) {
            float* pn = part + net * O::dnet;
            const int ho = net * HS, oo = net * OS, r0 = net * rotn;
            c_dw_gemm<MT, NW, NH, NO, false>(S, L.hb2 + ho, nullptr, L.go + oo, L.cst, pn + O::dw3, first, warp, lane, r0 % NW, exp);
            c_dw_gemm<MT, NW, NH, NH, false>(S, L.hb1 + ho, nullptr, L.gb2 + ho, L.cst, pn + O::dw2, first, warp, lane, (r0 + rot2) % NW, exp);
            c_dw_gemm<MT, NW, KS1, NH, true>(S, L.xt, nd->in_col, L.gb1 + ho, L.cst, pn + O::dw1, first, warp, lane, (r0 + rot3) % NW, exp);
        }
        c_stamp(exp, warp, lane);
        m_cta_sync();
    } else {
#pragma unroll 1
        for (int net = 0; net < 2; ++net) {
            uint32_t dout[NO][MT][4];
#pragma unroll
            for (int j = 0; j < NO; ++j)
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) dout[j][i][e] = net ? dt[j][i][e] : ds[j][i][e];
            if (net == 1) {   // recompute h1, h2 of the t subnet into the buffers
                float unused[NO][MT][4];
                c_subnet_keep<MT, TM, KS1, NH, NO, BD, false>(a1, Wn + O::net, lane, S + L.hb1, S + L.hb2, rp, unused);
            }
            c_store_afrag<MT, TM, NO>(S + L.go, rp, dout);
            c_dgrad<MT, NW, KS1, NH, NO, BD>(dout, Wt + net * O::tnet, nd->in_col, DZ, S + L.hb1, S + L.hb2, S + L.gb1, S + L.gb2, rp, warp, lane);
            c_stamp(exp, warp, lane);
            m_cta_sync();
            c_stamp(exp, warp, lane);
            float* pn = part + net * O::dnet;
            c_dw_gemm<MT, NW, NH, NO, false>(S, L.hb2, nullptr, L.go, L.cst, pn + O::dw3, first, warp, lane, 0, exp);
            c_dw_gemm<MT, NW, NH, NH, false>(S, L.hb1, nullptr, L.gb2, L.cst, pn + O::dw2, first, warp, lane, rot2, exp);
            c_dw_gemm<MT, NW, KS1, NH, true>(S, L.xt, nd->in_col, L.gb1, L.cst, pn + O::dw1, first, warp, lane, rot3, exp);
            c_stamp(exp, warp, lane);
            m_cta_sync();
        }
    }
}

template <int MT, int NW>
HINT_DEV void c_node_bwd_dispatch(const ChainNode* nd, float alpha, const float* __restrict__ W, float* S, const ChainBwdSmem& L,
                                  float* __restrict__ partial, bool first, int warp, int lane, int exp) {
    const float* Wn = W + (HINT_EXP(exp, 4) ? 0 : nd->w_off);
    const float* Wt = W + (HINT_EXP(exp, 4) ? 0 : nd->wt_off);
    float* part = partial + nd->dw_off;
#define HINT_CHAIN_CASE(ID, A, B, C, D) \
    case ID: c_node_bwd<MT, NW, A, B, C, D>(nd, alpha, Wn, Wt, S, L, part, first, warp, lane, exp); break;
    switch (nd->shape) {
        HINT_CHAIN_SHAPES(HINT_CHAIN_CASE)
        default: break;
    }
#undef HINT_CHAIN_CASE
}

// CTA-wide tile I/O (all threads): global rows [row0, row0+rows) x width <-> swizzled columns col_base ..
template <int TM, int NT>
HINT_DEV void c_load_tile_sw(float* buf, int col_base, const float* __restrict__ gsrc, long long row0, int rows, int width, int tid) {
    if (width == 0) return;
    const float* src = gsrc + row0 * width;
    const int nvalid = rows * width, n = TM * width;
    // (row, column) of element i without a division per iteration: one divmod up front, then constant strides
    int m0 = (tid * 4) / width, j0 = tid * 4 - m0 * width;
    const int dm = (NT * 4) / width, dj = NT * 4 - dm * width;
    for (int i = tid * 4; i < n; i += NT * 4) {
        float v[4];
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(src + i), "l"(c_policy_evict_first()));
#else
            for (int e = 0; e < 4; ++e) v[e] = src[i + e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (i + e < nvalid) ? src[i + e] : 0.f;
        }
        int m = m0, j = j0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            buf[c_elem<TM>(col_base + j, m)] = v[e];
            if (++j == width) { j = 0; ++m; }
        }
        m0 += dm; j0 += dj;
        if (j0 >= width) { j0 -= width; ++m0; }
    }
}
template <int TM, int NT>
HINT_DEV void c_store_tile_sw(const float* buf, int col_base, float* __restrict__ gdst, long long row0, int rows, int width, int tid) {
    if (width == 0) return;
    float* dst = gdst + row0 * width;
    const int nvalid = rows * width;
    int m0 = (tid * 4) / width, j0 = tid * 4 - m0 * width;
    const int dm = (NT * 4) / width, dj = NT * 4 - dm * width;
    for (int i = tid * 4; i < nvalid; i += NT * 4) {
        float v[4];
        int m = m0, j = j0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = buf[c_elem<TM>(col_base + j, m)];
            if (++j == width) { j = 0; ++m; }
        }
        m0 += dm; j0 += dj;
        if (j0 >= width) { j0 -= width; ++m0; }
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                         :: "l"(dst + i), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "l"(c_policy_evict_first()) : "memory");
#else
            for (int e = 0; e < 4; ++e) dst[i + e] = v[e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < nvalid) dst[i + e] = v[e];
        }
    }
}

template <int MT, int NW>
HINT_DEV void c_bwd_body(const ChainTables& T, const ChainNode* nodes, const ChainBwdSmem& L, float* S, const float* __restrict__ z,
                         const float* __restrict__ c, const float* __restrict__ W, const float* __restrict__ dz,
                         const float* __restrict__ dlogdet, float* __restrict__ x_rec, float* __restrict__ dx, float* __restrict__ dc,
                         float* __restrict__ partials, long long n_partial, long long B, int tid, int bid, int nblocks) {
    constexpr int TM = 16 * MT * NW, NT = 32 * NW;
    const int warp = tid >> 5, lane = tid & 31;
    float* partial = partials + (long long)bid * n_partial;
    const long long ntiles = (B + TM - 1) / TM;
    bool first = true;
    if (tid < 8) S[L.cst + tid] = tid < 4 ? 1.f : 0.f;
    for (long long tile = bid; tile < ntiles; tile += nblocks) {
        const long long row0 = tile * TM;
        const int rows = (int)((B - row0) < TM ? (B - row0) : TM);
        c_load_tile_sw<TM, NT>(S + L.xt, 0, z, row0, rows, T.d, tid);
        c_load_tile_sw<TM, NT>(S + L.xt, T.d, c, row0, rows, T.dc, tid);
        if (dz) c_load_tile_sw<TM, NT>(S + L.dz, 0, dz, row0, rows, T.d, tid);
        for (int i = tid; i < T.dc * TM; i += NT) S[L.dz + (T.d + i / TM) * TM + (i % TM)] = 0.f;
        for (int i = tid; i < TM; i += NT) S[L.dj + i] = (i < rows) ? (dlogdet ? dlogdet[row0 + i] : -T.nll_scale) : 0.f;
        m_cta_sync();
        if (!dz) {   // fused NLL gradient: the z tile (same swizzled layout) scaled into the gradient tile; rows past the batch are 0
            for (int i = tid; i < T.d * TM; i += NT) S[L.dz + i] = T.nll_scale * S[L.xt + i];
            m_cta_sync();
        }
        for (int q = T.n_nodes - 1; q >= 0; --q)
            c_node_bwd_dispatch<MT, NW>(nodes + q, T.alpha, W, S, L, partial, first, warp, lane, T.exp);
        if (x_rec) c_store_tile_sw<TM, NT>(S + L.xt, 0, x_rec, row0, rows, T.d, tid);
        c_store_tile_sw<TM, NT>(S + L.dz, 0, dx, row0, rows, T.d, tid);
        if (dc) c_store_tile_sw<TM, NT>(S + L.dz, T.d, dc, row0, rows, T.dc, tid);
        m_cta_sync();
        first = false;
    }
}

#if defined(__CUDACC__)
template <int MT, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
hint_bwd_chain_kernel(const __grid_constant__ ChainTables T, const __grid_constant__ ChainParam P, const __grid_constant__ ChainBwdSmem L,
                      const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ W,
                      const float* __restrict__ dz, const float* __restrict__ dlogdet, float* __restrict__ x_rec,
                      float* __restrict__ dx, float* __restrict__ dc, float* __restrict__ partials, long long n_partial, long long B) {
    extern __shared__ float4 c_smem4[];
    float* S = reinterpret_cast<float*>(c_smem4);
    int* nodes = reinterpret_cast<int*>(S + L.nodes);
    for (int i = threadIdx.x; i < T.n_nodes * 32; i += 32 * NW) nodes[i] = reinterpret_cast<const int*>(P.nodes)[i];
    __syncthreads();
    c_bwd_body<MT, NW>(T, reinterpret_cast<const ChainNode*>(nodes), L, S, z, c, W, dz, dlogdet, x_rec, dx, dc, partials, n_partial, B,
                       threadIdx.x, blockIdx.x, gridDim.x);
}
#endif

}  // namespace hint
