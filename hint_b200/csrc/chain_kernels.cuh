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
namespace hint { namespace emu { void warp_sync(); } }
#endif

namespace hint {

struct ChainTables {
    int n_nodes, d, dc;
    float alpha;
    int n_fwd_packed;   // floats of the forward operand region (staged in shared memory by the WS kernels)
};

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

// d = a*b + c with c in its own registers (the bias fragment is shared by all m-tiles: no accumulator initialisation moves)
HINT_DEV void c_mma_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
#if defined(__CUDA_ARCH__)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
#else
    d[0] = c0; d[1] = c1; d[2] = c0; d[3] = c1;
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
        float b0 = 0.f, b1 = 0.f, w0, w1;
        if (bl != nullptr) c_ldw2<WS>(bl + 8 * j + 2 * t, b0, b1);
        c_ldw2<WS>(Wl + j * 64 + 2 * lane, w0, w1);
#pragma unroll
        for (int i = 0; i < MT; ++i) c_mma_c(acc[j][i], a[0][i], m_bits(w0), m_bits(w1), b0, b1);
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

// layer-1 A fragments from the warp's x tile: input feature f < k is x column lo+f, k <= f < cin the condition column
// d + (f - k), beyond that zero (the packed W1 rows are zero there too)
template <int MT, int KS1>
HINT_DEV void c_load_input(const float* XT, int lo, int k, int cin, int d, int lane, uint32_t (&a)[KS1][MT][4]) {
    constexpr int PW = 16 * MT + 4, R = 2 * MT;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int f = 8 * ks + 2 * t + c;
            float v[R];
#pragma unroll
            for (int e = 0; e < R; ++e) v[e] = 0.f;
            if (f < cin) c_ld_rows<MT>(XT + (f < k ? lo + f : d + (f - k)) * PW + R * g, v);
#pragma unroll
            for (int i = 0; i < MT; ++i) { a[ks][i][2 * c] = m_rna(v[2 * i]); a[ks][i][2 * c + 1] = m_rna(v[2 * i + 1]); }
        }
}

// s / t subnet of one node (hint.py:10-13,77) entirely in registers
template <bool WS, int MT, int KS1, int NH, int NO>
HINT_DEV void c_subnet(const uint32_t (&a1)[KS1][MT][4], const float* __restrict__ Wnet, int lane, float (&out)[NO][MT][4]) {
    uint32_t h[NH][MT][4];
    {
        float acc[NH][MT][4];
        c_layer<WS, MT, KS1, NH>(a1, Wnet + chain_w1(KS1, NH, NO), Wnet + chain_b1(KS1, NH, NO), lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    {
        float acc[NH][MT][4];
        c_layer<WS, MT, NH, NH>(h, Wnet + chain_w2(KS1, NH, NO), Wnet + chain_b2(KS1, NH, NO), lane, acc);
        c_relu_to_a<MT, NH>(acc, h);
    }
    c_layer<WS, MT, NH, NO>(h, Wnet + chain_w3(KS1, NH, NO), Wnet + chain_b3(KS1, NH, NO), lane, out);
}

// one tree node, forward (hint.py:79-81) or inverse (hint.py:82-84) coupling; JP = the lane's private log-det partials
template <bool WS, int MT, int KS1, int NH, int NO, bool REV>
HINT_DEV void c_node_fwd(int lo, int k, int cout, int cin, int d, float alpha, const float* __restrict__ Wn, float* XT, float* JP,
                         int lane) {
    constexpr int PW = 16 * MT + 4, R = 2 * MT;
    const int g = lane >> 2, t = lane & 3;
    uint32_t a1[KS1][MT][4];
    c_load_input<MT, KS1>(XT, lo, k, cin, d, lane, a1);
    float s[NO][MT][4], tt[NO][MT][4];
    c_subnet<WS, MT, KS1, NH, NO>(a1, Wn, lane, s);
    c_subnet<WS, MT, KS1, NH, NO>(a1, Wn + chain_net_floats(KS1, NH, NO), lane, tt);
    float jl[R];
    c_ld_rows<MT>(JP + (t * 16 * MT) + R * g, jl);
#pragma unroll
    for (int j = 0; j < NO; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = 8 * j + 2 * t + c;
            if (col < cout) {
                float* xp = XT + (lo + k + col) * PW + R * g;
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
    c_st_rows<MT>(JP + (t * 16 * MT) + R * g, jl);
}

template <bool WS, int MT, bool REV>
HINT_DEV void c_node_fwd_dispatch(const ChainNode& nd, int d, float alpha, const float* __restrict__ W, float* XT, float* JP, int lane) {
    const float* Wn = W + nd.w_off;
#define HINT_CHAIN_CASE(ID, A, B, C) \
    case ID: c_node_fwd<WS, MT, A, B, C, REV>(nd.lo, nd.k, nd.cout, nd.cin, d, alpha, Wn, XT, JP, lane); break;
    switch (nd.shape) {
        HINT_CHAIN_CASE(0, 1, 1, 1) HINT_CHAIN_CASE(1, 1, 2, 1) HINT_CHAIN_CASE(2, 1, 3, 1) HINT_CHAIN_CASE(3, 1, 5, 1)
        HINT_CHAIN_CASE(4, 2, 5, 2) HINT_CHAIN_CASE(5, 2, 9, 2) HINT_CHAIN_CASE(6, 3, 9, 3)
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
    for (int i = lane * 4; i < n; i += 128) {
        float v[4];
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            // streamed once: keep the tile out of L1, which holds the weights
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(src + i));
#else
            for (int e = 0; e < 4; ++e) v[e] = src[i + e];
#endif
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (i + e < nvalid) ? src[i + e] : 0.f;
        }
        int m = i / width, j = i - m * width;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            XT[(col_base + j) * PW + m] = v[e];
            if (++j == width) { j = 0; ++m; }
        }
    }
}
template <int MT>
HINT_DEV void c_store_tile(const float* XT, int col_base, float* __restrict__ gdst, long long row0, int rows, int width, int lane) {
    constexpr int PW = 16 * MT + 4;
    if (width == 0) return;
    float* dst = gdst + row0 * width;
    const int nvalid = rows * width;
    for (int i = lane * 4; i < nvalid; i += 128) {
        float v[4];
        int m = i / width, j = i - m * width;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = XT[(col_base + j) * PW + m];
            if (++j == width) { j = 0; ++m; }
        }
        if (i + 3 < nvalid) {
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
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

// floats of one warp's private shared-memory region (x + condition columns, 4 log-det partial rows)
template <int MT>
HINT_HD constexpr int chain_fwd_warp_floats(int d, int dc) { return (d + dc) * (16 * MT + 4) + 4 * 16 * MT; }

template <int MT, int NW, bool REV, bool WS>
HINT_DEV void c_fwd_body(const ChainTables& T, const ChainNode* nodes, float* S, const float* __restrict__ x, const float* __restrict__ c,
                         const float* __restrict__ W, float* __restrict__ z, float* __restrict__ logdet, long long B, int tid,
                         int bid, int nblocks) {
    constexpr int RW = 16 * MT;
    const int warp = tid >> 5, lane = tid & 31;
    float* XT = S + warp * chain_fwd_warp_floats<MT>(T.d, T.dc);
    float* JP = XT + (T.d + T.dc) * (RW + 4);
    const long long ntiles = (B + RW - 1) / RW;
    for (long long tile = (long long)bid * NW + warp; tile < ntiles; tile += (long long)nblocks * NW) {
        const long long row0 = tile * RW;
        const int rows = (int)((B - row0) < RW ? (B - row0) : RW);
        c_load_tile<MT>(XT, 0, x, row0, rows, T.d, lane);
        c_load_tile<MT>(XT, T.d, c, row0, rows, T.dc, lane);
        for (int i = lane; i < 4 * RW; i += 32) JP[i] = 0.f;
        c_syncwarp();
        for (int q = 0; q < T.n_nodes; ++q) {
            const ChainNode& nd = nodes[REV ? T.n_nodes - 1 - q : q];
            c_node_fwd_dispatch<WS, MT, REV>(nd, T.d, T.alpha, W, XT, JP, lane);
            c_syncwarp();
        }
        c_store_tile<MT>(XT, 0, z, row0, rows, T.d, lane);
        if (lane < rows) logdet[row0 + lane] = JP[lane] + JP[RW + lane] + JP[2 * RW + lane] + JP[3 * RW + lane];
        c_syncwarp();
    }
}

// shared memory of the forward kernels: [node table | forward operands (WS only) | NW warp tiles]
HINT_HD constexpr int chain_node_floats(int n_nodes) { return n_nodes * 8; }
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
    for (int i = threadIdx.x; i < T.n_nodes * 8; i += 32 * NW) nodes[i] = reinterpret_cast<const int*>(P.nodes)[i];
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

}  // namespace hint
