// tcgen05 / TMEM TRAINING kernel: memory-free backward of one HINT coupling block (hint.py:62-101 and the autograd tape
// over it) in TF32.  Executes the programs of plan_tc3.h; the machine model is described there.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles of 128 samples):
//   warps 0-15 epilogue.  Thread = one sample (TMEM lane = tid & 127); the four warpgroups split the columns of the wide
//              steps (hidden layers, flushes) - with 2 warps per scheduler the epilogues were issue-latency bound; the
//              narrow, state-touching steps run on warpgroup 0 only, so the per-sample state in shared memory is private to
//              one thread and needs no barrier.
//   warp 16    MMA issuer (one elected thread).  All MMA operands derive from the kernel-parameter bank (program passed
//              by value) and uniform arithmetic, so tcgen05.mma issues without the compiler's divergence loop.
//   warp 17    loader: weight slabs -> shared-memory ring (cp.async.bulk, mbarrier complete_tx).
// Synchronisation: two monotone signal sequences on mbarrier rings (MMA groups committed, epilogue steps finished) with
// the wait indices the planner inferred; full/empty barriers per ring slot.
#pragma once
#include "plan_tc3.h"
#include "tcgen05.cuh"

namespace hint {

constexpr int kT3MaxMma = 1200;

// One issuer record as five 32-bit words in the kernel-parameter bank.  Sub-word fields are decoded with shifts, never with
// 16-bit struct members: ptxas 12.9 miscompiled `mov.b32 {%rs, %rs}` unpacks of dynamically indexed ld.param words (the
// consumer read a register the load never wrote), which made the issuer skip its waits.
//   w0 idesc | w1 b_off | w2 d_col | a_col << 16 | w3 nk | b_sbo16 << 16 | w4 flags | wait_epi << 16
struct T3MmaWords { uint32_t w[5]; };
inline T3MmaWords t3_pack_mma(const T3Mma& m) {
    T3MmaWords r;
    r.w[0] = m.idesc; r.w[1] = m.b_off;
    r.w[2] = (uint32_t)m.d_col | ((uint32_t)m.a_col << 16);
    r.w[3] = (uint32_t)m.nk | ((uint32_t)m.b_sbo16 << 16);
    r.w[4] = (uint32_t)m.flags | ((uint32_t)(uint16_t)m.wait_epi << 16);
    return r;
}

struct T3Prog {
    int n_mma, n_epi, n_chunks, n_signals;
    int n_slots, slot_bytes;
    int sm_bars, sm_tab16, sm_epis, sm_xs, sm_gs, sm_os, sm_red, sm_ring;
    int sm_img[kT3Imgs];
    int img_rows[kT3Imgs];
    int xp, op, d, dc, n_tab16;
    float alpha;
    int kind;          // T3K_BACKWARD, or T3K_FORWARD / T3K_INVERSE: the transport alone (z = in, dx = out, x_rec = log-det out)
    float nll_scale;   // used when dz == NULL: dz = nll_scale * z, dlogdet = -nll_scale (fused NLL gradient, train_unconditional.py:128-132)
    const T3Epi* epis;
    const T3Chunk* chunks;
    const int16_t* tab16;
    T3MmaWords mmas[kT3MaxMma];
};

// barrier slots (uint64_t) at sm_bars
enum { T3B_FULL = 0, T3B_EMPTY = 4, T3B_MMA = 8, T3B_EPI = 8 + kT3NB, T3B_DONE = 8 + 2 * kT3NB, T3B_COUNT = 9 + 2 * kT3NB };

__device__ __forceinline__ float t3_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
// atan(x): reciprocal range reduction + degree-7 minimax polynomial in x^2 (max abs error 1.7e-7)
__device__ __forceinline__ float t3_atan(float x) {
    const float a = fabsf(x);
    const bool inv = a > 1.f;
    float r = a;
    if (inv) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
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
// round to nearest (ties away) tf32, sign-magnitude arithmetic on the bit pattern
__device__ __forceinline__ float t3_round(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ uint64_t t3_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kT3Threads, 1)
hint_tc3_bwd_kernel(const __grid_constant__ T3Prog P, const float* __restrict__ z, const float* __restrict__ c,
                    const float* __restrict__ W, const float* __restrict__ dz, const float* __restrict__ dlogdet,
                    float* __restrict__ x_rec, float* __restrict__ dx, float* __restrict__ dcond, float* __restrict__ partials,
                    long long n_partial, long long B, float* dbg, long long* prof) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.sm_bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.sm_bars + 1000);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int16_t* s_tab = reinterpret_cast<int16_t*>(smem + P.sm_tab16);
    for (int i = tid; i < P.n_tab16; i += kT3Threads) s_tab[i] = P.tab16[i];
    {   // the epilogue program lives in shared memory: a global fetch per step was ~700 cycles of exposed L2 latency
        int4* s_epis4 = reinterpret_cast<int4*>(smem + P.sm_epis);
        const int4* g4 = reinterpret_cast<const int4*>(P.epis);
        for (int i = tid; i < P.n_epi * 3; i += kT3Threads) s_epis4[i] = g4[i];
    }
    if (tid == 0) {
        for (int i = 0; i < kT3MaxSlots; ++i) { mbar_init(bars + T3B_FULL + i, 1); mbar_init(bars + T3B_EMPTY + i, 1); }
        for (int i = 0; i < kT3NB; ++i) { mbar_init(bars + T3B_MMA + i, 1); mbar_init(bars + T3B_EPI + i, kT3EpiWarps); }
        mbar_init(bars + T3B_DONE, 1);
        fence_mbar_init();
    }
    if (warp == kT3EpiWarps) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (*tmem_slot != 0) __trap();   // the CTA owns the whole TMEM: base column 0 keeps the MMA operands uniform
    const long long ntiles = (B + 127) / 128;

    if (warp == kT3EpiWarps + 1) {
        // ================= loader =================
        if (elect_one()) {
            unsigned char* ring = smem + P.sm_ring;
            uint32_t slot = 0, ph = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < P.n_chunks; ++ch) {
                    // descriptor from the read-only path: the table is re-read for every tile and stays in L1.  A plain global
                    // load here cost an exposed L2 round trip per chunk and capped the ring at one chunk per ~600 cycles
                    // whatever its depth (every record of a hidden-layer GEMM waited for its slab).
                    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(P.chunks) + ch);
                    T3Chunk ck; ck.g_off = raw.x; ck.bytes = raw.y;
                    mbar_wait(bars + T3B_EMPTY + slot, ph ^ 1);
                    mbar_arrive_expect_tx(bars + T3B_FULL + slot, ck.bytes);
                    bulk_g2s(ring + (size_t)slot * P.slot_bytes, W + ck.g_off, ck.bytes, bars + T3B_FULL + slot);
                    if (++slot == (uint32_t)P.n_slots) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == kT3EpiWarps) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t sbase = smem_u32(smem);
            const uint32_t ring16 = (sbase + (uint32_t)P.sm_ring) >> 4;
            const uint32_t slot16 = (uint32_t)P.slot_bytes >> 4;
            uint32_t nslot = 0, nph = 0, cur = 0;
            uint32_t msig = 0;        // MMA signals committed so far (all tiles)
            uint32_t ebase = 0;       // epilogue steps of the previous tiles
            int tile_iter = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
                int waited = -1;
                const bool pt = blockIdx.x == 0 && tile_iter == 1;   // developer timeline of CTA 0's second tile
                // The records live in the kernel-parameter (constant) bank, which keeps every MMA operand uniform, but a tile walks
                // more records than the constant cache holds: fetching a record right before its use exposed a cache miss per
                // record (~350 cycles, the tensor pipe idling behind every 4-MMA record).  The NEXT record is therefore fetched while
                // the current one is being issued.
                uint32_t n0 = P.mmas[0].w[0], n1 = P.mmas[0].w[1], n2 = P.mmas[0].w[2], n3 = P.mmas[0].w[3], n4 = P.mmas[0].w[4];
                for (int i = 0; i < P.n_mma; ++i) {
                    const uint32_t w0 = n0, w1 = n1, w2 = n2, w3 = n3, w4 = n4;
                    {
                        const int nx = i + 1 < P.n_mma ? i + 1 : 0;
                        n0 = P.mmas[nx].w[0]; n1 = P.mmas[nx].w[1]; n2 = P.mmas[nx].w[2]; n3 = P.mmas[nx].w[3]; n4 = P.mmas[nx].w[4];
                    }
                    const uint32_t m_idesc = w0, m_boff = w1, m_dcol = w2 & 0xFFFFu, m_acol = w2 >> 16, m_nk = w3 & 0xFFFFu, m_sbo16 = w3 >> 16;
                    const uint32_t m_flags = w4 & 0xFFFFu;
                    const int m_wait = (int)w4 >> 16;   // arithmetic shift: sign-extended 16-bit field
                    if (m_wait > waited) {
                        const uint32_t g = ebase + (uint32_t)m_wait;
                        mbar_wait(bars + T3B_EPI + (g % kT3NB), (g / kT3NB) & 1);
                        fence_after_sync();
                        waited = m_wait;
                    }
                    if (prof && pt && i < 1024) prof[i] = clock64();
                    long long t_full = 0, t_mma = 0;
                    uint32_t acc = (m_flags & T3M_ZERO) ? 0u : 1u;
                    if (!(m_flags & T3M_SS)) {
                        if (m_flags & T3M_NEWCHUNK) {
                            cur = nslot;
                            mbar_wait(bars + T3B_FULL + cur, nph);
                            if (++nslot == (uint32_t)P.n_slots) { nslot = 0; nph ^= 1; }
                        }
                        if (prof && pt) t_full = clock64();
                        uint32_t b_lo = ((ring16 + cur * slot16 + (m_boff >> 4)) & 0x3FFFu) | (8u << 16);   // LBO = 128 B
                        const uint32_t b_hi = m_sbo16 | (1u << 14);                                            // SBO, descriptor version 1
                        uint32_t a_t = m_acol;
                        for (uint32_t ks = 0; ks < m_nk; ++ks) {
                            mma_ts(m_dcol, a_t, ((uint64_t)b_hi << 32) | b_lo, m_idesc, acc);
                            b_lo += 16;   // two core matrices (256 B) along K
                            a_t += 8;
                            acc = 1u;
                        }
                        if (m_flags & T3M_ENDCHUNK) commit(bars + T3B_EMPTY + cur);
                    } else {
                        const int ai = (int)(m_boff & 0xFF), bi = (int)((m_boff >> 8) & 0xFF);
                        const uint32_t abase = sbase + (uint32_t)P.sm_img[ai] + m_acol * 1024u;
                        const uint32_t bbase = sbase + (uint32_t)P.sm_img[bi];
                        const uint32_t aslab = (uint32_t)P.img_rows[ai] * 128u, bslab = (uint32_t)P.img_rows[bi] * 128u;
                        for (int kk = 0; kk < 16; ++kk) {
                            const uint32_t o = (uint32_t)(kk & 3) * 32u;
                            mma_ss(m_dcol, t3_desc_sw128(abase + (uint32_t)(kk >> 2) * aslab + o),
                                   t3_desc_sw128(bbase + (uint32_t)(kk >> 2) * bslab + o), m_idesc, acc);
                            acc = 1u;
                        }
                    }
                    if (prof && pt) t_mma = clock64();
                    if (m_flags & T3M_COMMIT) { commit(bars + T3B_MMA + (msig % kT3NB)); ++msig; }
                    if (prof && pt && i < 512) { prof[4096 + 3 * i] = t_full; prof[4097 + 3 * i] = t_mma; prof[4098 + 3 * i] = clock64(); }
                }
                ebase += (uint32_t)P.n_epi;
                if (dbg) commit(bars + T3B_DONE);   // developer dump (tests/cuda/dbg_tc3.py): everything issued has completed
            }
        }
    } else {
        // ================= epilogue warps =================
        const int wg = tid >> 7;            // warpgroup 0 .. 3: column slice of the wide steps
        constexpr int kWG = kT3EpiWarps / 4;
        const int row = tid & 127;          // sample within the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float* XS = reinterpret_cast<float*>(smem + P.sm_xs) + row * P.xp;
        float* GS = reinterpret_cast<float*>(smem + P.sm_gs) + row * P.xp;
        float* OS = reinterpret_cast<float*>(smem + P.sm_os) + row * P.op;
        float* s_red = reinterpret_cast<float*>(smem + P.sm_red);
        const int d = P.d, dc = P.dc, nd = d + dc;
        const float alpha = P.alpha;
        // SWIZZLE_128B image addressing of this thread's sample: byte offset inside an 8-row group, per row & 7
        uint32_t xo[8];
#pragma unroll
        for (int e7 = 0; e7 < 8; ++e7) xo[e7] = (uint32_t)e7 * 128u + (uint32_t)((((row & 31) >> 2) ^ e7) << 4) + (uint32_t)(row & 3) * 4u;
        const uint32_t slab_q = (uint32_t)(row >> 5) * 128u;   // x img_rows = byte offset of this sample's 32-sample slab
        auto img_ptr = [&](int i) -> unsigned char* { return smem + P.sm_img[i] + slab_q * (uint32_t)P.img_rows[i]; };
        float* my_part = partials + (size_t)blockIdx.x * (size_t)n_partial;
        uint32_t sbase_sig = 0;   // MMA signals of the previous tiles
        uint32_t estep = 0;       // epilogue steps completed so far (all tiles)
        bool first_tile = true;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long row0 = tile * 128;
            const int rows = (int)((B - row0) < 128 ? (B - row0) : 128);
            // ---- tile state: [z | c], [dz | 0], dJ ----
            {
                float* xs_all = reinterpret_cast<float*>(smem + P.sm_xs);
                float* gs_all = reinterpret_cast<float*>(smem + P.sm_gs);
                const int nx = rows * d;
                const float* gz = z + row0 * d;
                const float* gdz = dz + row0 * d;
                if (P.kind != T3K_BACKWARD) {
                    for (int i = tid; i < 128 * d; i += 32 * kT3EpiWarps) {
                        const int s = i / d, j = i - s * d;
                        xs_all[s * P.xp + j] = i < nx ? __ldg(gz + i) : 0.f;
                    }
                } else {
                    for (int i = tid; i < 128 * d; i += 32 * kT3EpiWarps) {
                        const int s = i / d, j = i - s * d;
                        xs_all[s * P.xp + j] = i < nx ? __ldg(gz + i) : 0.f;
                        gs_all[s * P.xp + j] = i < nx ? (dz ? __ldg(gdz + i) : P.nll_scale * __ldg(gz + i)) : 0.f;
                    }
                }
                if (dc) {
                    const int nc = rows * dc;
                    const float* gc = c + row0 * dc;
                    for (int i = tid; i < 128 * dc; i += 32 * kT3EpiWarps) {
                        const int s = i / dc, j = i - s * dc;
                        xs_all[s * P.xp + d + j] = i < nc ? __ldg(gc + i) : 0.f;
                        if (P.kind == T3K_BACKWARD) gs_all[s * P.xp + d + j] = 0.f;   // the transport programs have no gradient state
                    }
                }
            }
            const float dJ = (P.kind == T3K_BACKWARD && row < rows) ? (dlogdet ? __ldg(dlogdet + row0 + row) : -P.nll_scale) : 0.f;
            float Jacc = 0.f;   // transport kernels: log|det J| of this thread's sample (warpgroup 0)
            named_bar_sync(1, 32 * kT3EpiWarps);
            int waited = -1;
            for (int si = 0; si < P.n_epi; ++si, ++estep) {
                T3Epi e;
                {
                    const int4* ep = reinterpret_cast<const int4*>(smem + P.sm_epis) + 3 * si;
                    const int4 e0 = ep[0], e1 = ep[1], e2 = ep[2];
                    e.type = e0.x; e.flags = e0.y; e.wait_mma = e0.z; e.off = e0.w;
                    e.a = e1.x; e.b = e1.y; e.c = e1.z; e.d = e1.w; e.e = e2.x; e.f = e2.y; e.g = e2.z; e.h = e2.w;
                }
                if ((int)e.wait_mma > waited) {
                    const uint32_t g = sbase_sig + (uint32_t)e.wait_mma;
                    mbar_wait(bars + T3B_MMA + (g % kT3NB), (g / kT3NB) & 1);
                    waited = e.wait_mma;
                }
                long long t_wait_done = 0;
                if (prof) t_wait_done = clock64();
                fence_after_sync();
                switch (e.type) {
                    case T3E_IN: {
                        if (wg == 0) {
                            unsigned char* im3 = img_ptr(3);
                            const bool to_img = !(e.flags & T3I_NOIMG);
                            for (int c0 = 0; c0 < e.b; c0 += 8) {
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int code = s_tab[e.a + c0 + j];
                                    v[j] = t3_round(code >= 0 ? XS[code] : (code == -2 ? 1.f : 0.f));
                                    if (to_img) *reinterpret_cast<float*>(im3 + (uint32_t)(c0 >> 3) * 1024u + xo[j]) = v[j];
                                }
                                st8(lane_base + (uint32_t)(e.c + c0), v);
                            }
                        }
                        break;
                    }
                    case T3E_HID: {
                        const int ncols = e.b;
                        const int slice = ((ncols + kWG - 1) / kWG + 15) & ~15;
                        const int q0 = min(wg * slice, ncols), q1 = min(q0 + slice, ncols);
                        const uint32_t a0 = lane_base + (uint32_t)e.a;
                        unsigned char* im = img_ptr(e.c);
                        const bool to_img = (e.flags & T3H_IMG) != 0;
                        // 16 columns: relu + tf32 rounding in place (A operand of the next GEMM) (+ image: operand of dW / relu mask)
                        auto finish16 = [&](int q, float (&v)[16]) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = t3_round(fmaxf(v[j], 0.f));
                            st16(a0 + q, v);
                            if (to_img) {
                                unsigned char* ib = im + (uint32_t)(q >> 3) * 1024u;
#pragma unroll
                                for (int j = 0; j < 16; ++j) *reinterpret_cast<float*>(ib + (uint32_t)(j >> 3) * 1024u + xo[j & 7]) = v[j];
                            }
                        };
                        int q = q0;
                        for (; q + 32 <= q1; q += 32) {       // two loads in flight
                            float v0[16], v1[16];
                            ld16(a0 + q, v0); ld16(a0 + q + 16, v1);
                            wait_ld();
                            finish16(q, v0);
                            finish16(q + 16, v1);
                        }
                        if (q < q1) {
                            float v0[16];
                            ld16(a0 + q, v0);
                            wait_ld();
                            finish16(q, v0);
                        }
                        if (wg == kWG - 1 && (e.flags & T3H_ONES)) {
                            float o[8] = {1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // two ones: bias hi + lo parts (kT3BiasLo)
                            st8(a0 + ncols, o);
                            if (e.flags & T3H_IMG_ONES) {
                                unsigned char* ib = im + (uint32_t)(ncols >> 3) * 1024u;
#pragma unroll
                                for (int j = 0; j < 8; ++j) *reinterpret_cast<float*>(ib + xo[j]) = o[j];
                            }
                        }
                        break;
                    }
                    case T3E_OUTS: {
                        if (wg == 0) {
                            for (int c0 = 0; c0 < e.b; c0 += 8) {
                                float v[8];
                                ld8(lane_base + (uint32_t)(e.a + c0), v);
                                wait_ld();
#pragma unroll
                                for (int j = 0; j < 8; ++j) if (c0 + j < e.b) OS[c0 + j] = v[j];
                            }
                        }
                        break;
                    }
                    case T3E_CPL:
                    case T3E_DS: {
                        if (wg == 0) {
                            const bool cpl = e.type == T3E_CPL;
                            const int oc = e.b, kd = e.e;
                            const int tm_dout = cpl ? e.d : e.a;
                            unsigned char* im4 = img_ptr(4);
                            float* pq = my_part + e.off + (warp & 3) * 32;   // this quadrant's bias-gradient slots
                            for (int c0 = 0; c0 < kd; c0 += 8) {
                                float tv[8], o[8], bs[8];
                                if (cpl) { ld8(lane_base + (uint32_t)(e.a + c0), tv); wait_ld(); }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int col = c0 + j;
                                    float dout = 0.f;
                                    if (col < oc) {
                                        if (cpl) {
                                            const int xc = s_tab[e.c + col];
                                            const float sv = OS[col];
                                            const float zl = XS[xc], dzl = GS[xc];
                                            const float la = alpha * t3_atan(sv);
                                            const float ee = t3_exp(la);
                                            const float xl = (zl - tv[j]) * t3_exp(-la);
                                            XS[xc] = xl;
                                            GS[xc] = dzl * ee;
                                            OS[col] = (dzl * xl * ee + dJ) * alpha / fmaf(sv, sv, 1.f);
                                            dout = dzl;
                                        } else {
                                            dout = OS[col];
                                        }
                                        *reinterpret_cast<float*>(im4 + (uint32_t)(c0 >> 3) * 1024u + xo[j]) = t3_round(dout);
                                    }
                                    bs[j] = dout;
                                    o[j] = t3_round(dout);
                                }
                                st8(lane_base + (uint32_t)(tm_dout + c0), o);
                                // bias gradient of layer 3: the 8 column sums over this quadrant's 32 samples as a reduce-scatter
                                // butterfly (9 shuffles instead of 40): afterwards lane 4*k holds the sum of column k
                                {
                                    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0, hi4 = (lane & 4) != 0;
                                    float a4[4], a2[2], a1;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float send = hi16 ? bs[k] : bs[k + 4];
                                        const float keep = hi16 ? bs[k + 4] : bs[k];
                                        a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                                    }
#pragma unroll
                                    for (int k = 0; k < 2; ++k) {
                                        const float send = hi8 ? a4[k] : a4[k + 2];
                                        const float keep = hi8 ? a4[k + 2] : a4[k];
                                        a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                                    }
                                    {
                                        const float send = hi4 ? a2[0] : a2[1];
                                        const float keep = hi4 ? a2[1] : a2[0];
                                        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                                    }
                                    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
                                    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                                    // lane bits (16, 8, 4) select the column: col = 4*bit16 + 2*bit8 + bit4
                                    const int kcol = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                                    if ((lane & 3) == 0 && c0 + kcol < oc) {
                                        float* p1 = pq + c0 + kcol;
                                        if (first_tile) *p1 = a1; else atomicAdd(p1, a1);   // private slot: plain accumulation, issued as RED (no round trip)
                                    }
                                }
                            }
                        }
                        break;
                    }
                    case T3E_CPLF: {   // the coupling of the transport kernels (hint.py:79-84): s saved by OUTS, t fresh in TMEM
                        if (wg == 0) {
                            const bool inv = e.flags != 0;
                            for (int c0 = 0; c0 < e.b; c0 += 8) {
                                float tv[8];
                                ld8(lane_base + (uint32_t)(e.a + c0), tv);
                                wait_ld();
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int col = c0 + j;
                                    if (col < e.b) {
                                        const int xc = s_tab[e.c + col];
                                        const float la = alpha * t3_atan(OS[col]);
                                        if (!inv) { XS[xc] = fmaf(t3_exp(la), XS[xc], tv[j]); Jacc += la; }
                                        else { XS[xc] = (XS[xc] - tv[j]) * t3_exp(-la); Jacc -= la; }
                                    }
                                }
                            }
                        }
                        break;
                    }
                    case T3E_DHID: {
                        const int ncols = e.b;
                        const int slice = ((ncols + kWG - 1) / kWG + 15) & ~15;
                        const int q0 = min(wg * slice, ncols), q1 = min(q0 + slice, ncols);
                        const uint32_t a0 = lane_base + (uint32_t)e.a;
                        unsigned char* imo = img_ptr(e.c);
                        // 16 columns: v = relu'(mask) ? round(v) : 0 -> TMEM (A operand of the next dgrad GEMM) + image (operand of dW)
                        auto finish16 = [&](int q, float (&v)[16], const float (&hm)[16]) {
                            unsigned char* ib = imo + (uint32_t)(q >> 3) * 1024u;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                v[j] = hm[j] > 0.f ? t3_round(v[j]) : 0.f;
                                *reinterpret_cast<float*>(ib + (uint32_t)(j >> 3) * 1024u + xo[j & 7]) = v[j];
                            }
                            st16(a0 + q, v);
                        };
                        if (e.flags & T3D_MASK_TMEM) {      // relu mask = the forward activation still in TMEM
                            const uint32_t m0 = lane_base + (uint32_t)e.e;
                            int q = q0;
                            for (; q + 32 <= q1; q += 32) {   // two loads in flight
                                float v0[16], v1[16], h0[16], h1[16];
                                ld16(a0 + q, v0); ld16(m0 + q, h0); ld16(a0 + q + 16, v1); ld16(m0 + q + 16, h1);
                                wait_ld();
                                finish16(q, v0, h0);
                                finish16(q + 16, v1, h1);
                            }
                            if (q < q1) {
                                float v0[16], h0[16];
                                ld16(a0 + q, v0); ld16(m0 + q, h0);
                                wait_ld();
                                finish16(q, v0, h0);
                            }
                        } else {                            // relu mask = the forward activation image in shared memory
                            const unsigned char* imm = img_ptr(e.e);
                            auto mask16 = [&](int q, float (&hm)[16]) {
                                const unsigned char* ib = imm + (uint32_t)(q >> 3) * 1024u;
#pragma unroll
                                for (int j = 0; j < 16; ++j) hm[j] = *reinterpret_cast<const float*>(ib + (uint32_t)(j >> 3) * 1024u + xo[j & 7]);
                            };
                            int q = q0;
                            for (; q + 32 <= q1; q += 32) {
                                float v0[16], v1[16], h0[16], h1[16];
                                ld16(a0 + q, v0); ld16(a0 + q + 16, v1);
                                mask16(q, h0); mask16(q + 16, h1);
                                wait_ld();
                                finish16(q, v0, h0);
                                finish16(q + 16, v1, h1);
                            }
                            if (q < q1) {
                                float v0[16], h0[16];
                                ld16(a0 + q, v0);
                                mask16(q, h0);
                                wait_ld();
                                finish16(q, v0, h0);
                            }
                        }
                        break;
                    }
                    case T3E_DA: {
                        if (wg == 0) {
                            for (int c0 = 0; c0 < e.b; c0 += 8) {
                                float v[8];
                                ld8(lane_base + (uint32_t)(e.a + c0), v);
                                wait_ld();
#pragma unroll
                                for (int j = 0; j < 8; ++j) if (c0 + j < e.b) GS[s_tab[e.c + c0 + j]] += v[j];
                            }
                        }
                        break;
                    }
                    case T3E_FLUSH: {
                        // accumulator rows = features on the lanes; this lane's node decides which column CHUNKS are real.  A chunk
                        // of 16 columns is flushed whole when it intersects the lane's ranges (node widths are padded to 16, so the
                        // hidden ranges are chunk-aligned; the few extra columns land in partial-buffer slots no parameter maps to).
                        int c0 = 0, c1 = 0;
                        bool found = false;
                        if (row < e.e) {
                            const int frow = e.h * 128 + row;
                            const int nn = s_tab[e.f];
                            for (int q = 0; q < nn; ++q) {
                                const int16_t* nt = s_tab + e.f + 1 + 6 * q;
                                if (frow >= nt[0] && frow < nt[0] + nt[1]) {
                                    found = true;
                                    const int k = e.g == T3F_W2 ? 0 : e.g == T3F_W1 ? 2 : 4;
                                    c0 = nt[k]; c1 = nt[k] + nt[k + 1];
                                }
                            }
                        }
                        const int x0 = e.c, x1 = e.g == T3F_W2 ? e.c + 1 : e.g == T3F_W1 ? e.c + dc + 1 : e.c;
                        float* pp = my_part + e.off + row;
                        const int ncols = e.b;
                        // accumulate into the CTA's private, L2-resident partial buffer.  Every address is owned by one thread and
                        // tiles are sequential, so a reduction without return value (RED: no L2 round trip on the critical path)
                        // is still a deterministic, uncontended accumulation; the first tile stores.
                        for (int q = 16 * wg; q < ncols; q += 16 * kWG) {
                            float v[16];
                            ld16(lane_base + (uint32_t)(e.a + q), v);
                            wait_ld();
                            const bool live = found && ((q < c1 && q + 16 > c0) || (x0 >= 0 && q < x1 && q + 16 > x0));
                            if (live) {
                                float* p0 = pp + (size_t)q * 128;
                                if (first_tile) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) __stcg(p0 + (size_t)j * 128, v[j]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) atomicAdd(p0 + (size_t)j * 128, v[j]);
                                }
                            }
                        }
                        break;
                    }
                    default: break;
                }
                // publish: TMEM / image writes of this step are visible to the tensor core before the signal fires
                long long ts0 = 0, ts1 = 0, ts2 = 0;
                if (prof) ts0 = clock64();
                wait_st();
                if (prof) ts1 = clock64();
                fence_proxy_async_smem();
                if (prof) ts2 = clock64();
                fence_before_sync();
                __syncwarp();
                if (prof && blockIdx.x == 0 && !first_tile && estep < 2u * (uint32_t)P.n_epi && si < 512 && lane == 0 && (warp == 0 || warp == 4)) {   // warpgroups 0 and 1
                    prof[1024 + 2 * si + (warp >> 2)] = clock64();
                    if (warp == 0) { prof[2048 + si] = t_wait_done; prof[2560 + 3 * si] = ts0; prof[2561 + 3 * si] = ts1; prof[2562 + 3 * si] = ts2; }
                }
                if (lane == 0) mbar_arrive(bars + T3B_EPI + (estep % kT3NB));
            }
            sbase_sig += (uint32_t)P.n_signals;
            if (dbg) {   // developer dump of one tile: TMEM [128][512] floats, then the raw dynamic shared memory
                mbar_wait(bars + T3B_DONE, 0);
                fence_after_sync();
                named_bar_sync(1, 32 * kT3EpiWarps);
                for (int q = (512 / kWG) * wg; q < (512 / kWG) * (wg + 1); q += 16) {
                    float v[16];
                    ld16(lane_base + (uint32_t)q, v);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) dbg[row * 512 + q + j] = v[j];
                }
                const int nwords = (int)(P.sm_red + 512) / 4;
                for (int i = tid; i < nwords; i += 32 * kT3EpiWarps) dbg[65536 + i] = reinterpret_cast<const float*>(smem)[i];
            }
            // ---- tile state -> dx, dc, x_rec ----
            named_bar_sync(1, 32 * kT3EpiWarps);
            {
                const float* xs_all = reinterpret_cast<const float*>(smem + P.sm_xs);
                const float* gs_all = reinterpret_cast<const float*>(smem + P.sm_gs);
                const int nx = rows * d;
                if (P.kind != T3K_BACKWARD) {
                    for (int i = tid; i < nx; i += 32 * kT3EpiWarps) {
                        const int s = i / d, j = i - s * d;
                        dx[row0 * d + i] = xs_all[s * P.xp + j];
                    }
                    if (wg == 0 && row < rows) x_rec[row0 + row] = Jacc;
                } else
                for (int i = tid; i < nx; i += 32 * kT3EpiWarps) {
                    const int s = i / d, j = i - s * d;
                    dx[row0 * d + i] = gs_all[s * P.xp + j];
                    if (x_rec) x_rec[row0 * d + i] = xs_all[s * P.xp + j];
                }
                if (dc && dcond && P.kind == T3K_BACKWARD) {
                    const int nc = rows * dc;
                    for (int i = tid; i < nc; i += 32 * kT3EpiWarps) {
                        const int s = i / dc, j = i - s * dc;
                        dcond[row0 * dc + i] = gs_all[s * P.xp + d + j];
                    }
                }
            }
            named_bar_sync(1, 32 * kT3EpiWarps);
            first_tile = false;
            (void)nd;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kT3EpiWarps) {
        __syncwarp();
        tmem_dealloc(0, 512);
    }
}

// dparams[dst[i]] = sum over the CTAs' partial buffers (fixed order: deterministic), reading the partials coalesced
__global__ void hint_tc3_reduce_kernel(const int* __restrict__ dst, const float* __restrict__ partials, int nctas, long long n_partial,
                                       float* __restrict__ dparams) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_partial; i += (long long)gridDim.x * blockDim.x) {
        const int t = dst[i];
        if (t < 0) continue;
        const float* p = partials + i;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (t & (1 << 30)) {   // layer-3 bias: four per-quadrant slots at stride 32
            for (int q = 0; q < nctas; ++q) {
                const float* pq = p + (long long)q * n_partial;
                a0 += pq[0]; a1 += pq[32]; a2 += pq[64]; a3 += pq[96];
            }
            dparams[t & ~(1 << 30)] = (a0 + a1) + (a2 + a3);
            continue;
        }
        int q = 0;
        for (; q + 3 < nctas; q += 4) {
            a0 += p[(long long)q * n_partial];
            a1 += p[(long long)(q + 1) * n_partial];
            a2 += p[(long long)(q + 2) * n_partial];
            a3 += p[(long long)(q + 3) * n_partial];
        }
        for (; q < nctas; ++q) a0 += p[(long long)q * n_partial];
        dparams[t] = (a0 + a1) + (a2 + a3);
    }
}

// packed[i] = tf32(params[src[i]]) (0 where src < 0)
__global__ void hint_tc3_pack_kernel(const int* __restrict__ src, const float* __restrict__ params, float* __restrict__ packed, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = src[i];
        if (s < 0) { packed[i] = 0.f; continue; }
        const float v = params[s & ~kT3BiasLo];
        const float hi = tc::to_tf32(v);
        packed[i] = (s & kT3BiasLo) ? tc::to_tf32(v - hi) : hi;
    }
}

}  // namespace hint
