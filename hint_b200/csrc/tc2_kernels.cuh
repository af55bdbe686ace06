// tcgen05 (TF32) fused-tree forward / inverse kernel, second generation.  Executes the program of plan_tc.h
// (T2Prog), which arrives BY VALUE in the kernel-parameter (constant) bank.
//
// What changed against the first tcgen05 kernel (round 1, removed), each item from a measurement (profiles/ubench3_r01_mma_issue_tmem.txt):
//   * MMA operands are warp-uniform (constant-bank loads + uniform arithmetic, TMEM base 0 because the CTA owns
//     all 512 columns): 38 cycles per tcgen05.mma instead of 77 behind the compiler's divergence loop;
//   * ops are pre-partitioned per issuing warp; an issuer never walks ops it does not own;
//   * epilogues move TMEM in batches of 4 x 16 columns per wait (43 elements/cycle/SM round trip against 12-26
//     with one wait per 16 columns);
//   * the next tile's rows are prefetched into shared memory by a bulk copy while the current tile computes;
//   * the coupling uses ex2-based exp and a multiply by exp(-log e) instead of a division.
//
// Warp roles (416 threads, one CTA per SM, persistent over tiles of 128 samples = 128 TMEM lanes):
//   warps 0-7   epilogue (warp w owns lanes 32*(w%4)..+31; warps 0-3 / 4-7 split the columns of every job)
//   warps 8-11  MMA issuers (one elected thread each)
//   warp 12     producer: weight chunks -> ring (bulk copies), x / c tiles -> staging (bulk copies)
#pragma once
#include "plan_tc.h"
#include "tcgen05.cuh"

namespace hint {

constexpr int kT2Threads = 256 + 32 * kTcIssuers + 32;
constexpr int kT2EpiThreads = 256;

// barrier block layout (uint64_t slots at the start of dynamic shared memory)
enum { T2B_FULL = 0, T2B_EMPTY = 8, T2B_XFULL = 16, T2B_XFREE = 18, T2B_TILE = 20, T2B_MMA = 21, T2B_EPI = 26, T2B_COUNT = 30 };

// exp(x) through ex2.approx (|x| <= 4 here: relative error ~2e-7)
__device__ __forceinline__ float t2_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
// atan(x): reciprocal range reduction to [0,1] + degree-7 minimax polynomial in x^2 (max abs error 1.7e-7, fitted in
// float32 arithmetic; the fit script is in DESIGN.md)
__device__ __forceinline__ float t2_atan(float x) {
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
// round-to-nearest (ties away) to tf32 for NON-NEGATIVE finite inputs with two integer ops; cvt.rna.tf32.f32 runs on the
// conversion unit at 16 lanes/cycle/SM and was 12k of the 23k cycles per tile of the hidden-layer epilogues
__device__ __forceinline__ float t2_round_tf32_pos(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

#define HINT_T2_T(var) do { if (dbg) { long long now_ = clock64(); (var) += now_ - tmark; tmark = now_; } } while (0)

template <bool kRev>
__global__ void __launch_bounds__(kT2Threads, 1)
hint_tc2_kernel(const __grid_constant__ T2Prog P, const float* __restrict__ x, const float* __restrict__ c,
                const float* __restrict__ W, float* __restrict__ z, float* __restrict__ logdet, long long B, long long* dbg) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 1016);
    float* s_bias = reinterpret_cast<float*>(smem + P.smem_bias);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nstages = P.nstages;
    const int n_slots = P.n_slots;
    // the epilogue warps read their tables from shared memory: walking them in the constant bank evicted the issuers' ops
    T2Stage* s_stages = reinterpret_cast<T2Stage*>(smem + P.smem_tab);
    T2Fin* s_fins = reinterpret_cast<T2Fin*>(s_stages + nstages);
    int16_t* s_xlog = reinterpret_cast<int16_t*>(s_fins + P.nfins);
    {
        const int n1 = nstages * (int)(sizeof(T2Stage) / 2), n2 = P.nfins * (int)(sizeof(T2Fin) / 2);
        const uint16_t* g1 = reinterpret_cast<const uint16_t*>(P.stages);
        const uint16_t* g2 = reinterpret_cast<const uint16_t*>(P.fins);
        for (int i = tid; i < n1; i += kT2Threads) reinterpret_cast<uint16_t*>(s_stages)[i] = g1[i];
        for (int i = tid; i < n2; i += kT2Threads) reinterpret_cast<uint16_t*>(s_fins)[i] = g2[i];
        for (int i = tid; i < P.xw; i += kT2Threads) s_xlog[i] = P.xlog[i];
    }
    for (int i = tid; i < P.n_bias; i += kT2Threads) s_bias[i] = W[P.bias_base + i];
    if (tid == 0) {
        for (int i = 0; i < n_slots; ++i) { mbar_init(bars + T2B_FULL + i, 1); mbar_init(bars + T2B_EMPTY + i, kTcIssuers); }
        for (int i = 0; i < 2; ++i) { mbar_init(bars + T2B_XFULL + i, 1); mbar_init(bars + T2B_XFREE + i, 8); }
        mbar_init(bars + T2B_TILE, 8);
        for (int i = 0; i < TC_NJOBS; ++i) mbar_init(bars + T2B_MMA + i, kTcIssuers);
        for (int i = 0; i < 4; ++i) mbar_init(bars + T2B_EPI + i, 8);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (*tmem_slot != 0) __trap();   // the CTA owns the whole TMEM, so the base is column 0 (keeps MMA operands uniform)
    const long long ntiles = (B + 127) / 128;
    const int x_bytes = 128 * P.d * 4, c_bytes = 128 * P.dc * 4;

    if (warp == 8 + kTcIssuers) {
        // ================= producer =================
        if (elect_one()) {
            unsigned char* ring = smem + P.smem_ring;
            uint32_t slot = 0, ph = 0, it = 0;
            auto prefetch = [&](long long tile, uint32_t k) {   // k-th tile of this CTA -> staging buffer k & 1
                const uint32_t buf = k & 1;
                if (k >= 2) mbar_wait(bars + T2B_XFREE + buf, ((k >> 1) - 1) & 1);
                float* dst = reinterpret_cast<float*>(smem + P.smem_in + buf * P.smem_in_bytes);
                if ((tile + 1) * 128 <= B) {
                    mbar_arrive_expect_tx(bars + T2B_XFULL + buf, (uint32_t)(x_bytes + c_bytes));
                    bulk_g2s(dst, x + tile * 128 * P.d, (uint32_t)x_bytes, bars + T2B_XFULL + buf);
                    if (P.dc) bulk_g2s(dst + 128 * P.d, c + tile * 128 * P.dc, (uint32_t)c_bytes, bars + T2B_XFULL + buf);
                } else {
                    mbar_arrive(bars + T2B_XFULL + buf);   // ragged last tile: the epilogue warps read global memory themselves
                }
            };
            if ((long long)blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x, it + 1);
                for (int si = 0; si < nstages; ++si) {
                    const T2Stage& S = P.stages[kRev ? si : nstages - 1 - si];
                    for (int ch = S.chunk_begin; ch < S.chunk_end; ++ch) {
                        mbar_wait(bars + T2B_EMPTY + slot, ph ^ 1);
                        const T2Chunk ck = P.chunks[ch];
                        mbar_arrive_expect_tx(bars + T2B_FULL + slot, ck.bytes);
                        bulk_g2s(ring + (size_t)slot * P.slot_bytes, W + (size_t)ck.g_off16 * 4, ck.bytes, bars + T2B_FULL + slot);
                        if (++slot == (uint32_t)n_slots) { slot = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ================= MMA issuers =================
        const int me = warp - 8;                       // warp-uniform
        if (elect_one()) {
            const uint32_t ring16 = (smem_u32(smem) + (uint32_t)P.smem_ring) >> 4;
            const uint32_t slot16 = (uint32_t)P.slot_bytes >> 4;
            uint32_t nslot = 0, nph = 0, it = 0, sctr = 0;   // next ring slot to consume and its phase
            long long t_tile = 0, t_prev = 0, t_epi = 0, t_chunk = 0, t_issue = 0, tmark = clock64();
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                mbar_wait(bars + T2B_TILE, it & 1);
                fence_after_sync();
                HINT_T2_T(t_tile);
                for (int si = 0; si < nstages; ++si, ++sctr) {
                    const T2Stage& S = P.stages[kRev ? si : nstages - 1 - si];
                    const uint32_t par = sctr & 1;
                    if (si > 0) {   // x columns written by the previous stage's coupling epilogue
                        mbar_wait(bars + T2B_EPI + 3, par ^ 1);
                        fence_after_sync();
                        HINT_T2_T(t_prev);
                    }
                    uint32_t slot = 0, waited = 0;
                    for (int g = S.seg_begin; g < S.seg_end; ++g) {
                        const T2Seg& G = P.segs[g];
                        const uint32_t job = G.job, fl = G.flags;
                        if (fl & T2_FIRST_IN_JOB) {
                            const int dep = job == TC_J1 ? -1 : (job == TC_J2S || job == TC_J2T) ? 0 : (job == TC_J3S ? 1 : 2);
                            if (dep >= 0 && !((waited >> dep) & 1)) {
                                mbar_wait(bars + T2B_EPI + dep, par);
                                fence_after_sync();
                                waited |= 1u << dep;
                                HINT_T2_T(t_epi);
                            }
                        }
                        if (fl & T2_FIRST_IN_CHUNK) {
                            slot = nslot;
                            mbar_wait(bars + T2B_FULL + slot, nph);
                            if (++nslot == (uint32_t)n_slots) { nslot = 0; nph ^= 1; }
                            HINT_T2_T(t_chunk);
                        }
                        const uint32_t base16 = ring16 + slot * slot16;
                        const int o0 = G.op_ofs[me], o1 = G.op_ofs[me + 1];
                        T2Op nxt = P.ops[o0 < o1 ? o0 : 0];
                        for (int oi = o0; oi < o1; ++oi) {
                            const T2Op op = nxt;
                            if (oi + 1 < o1) nxt = P.ops[oi + 1];   // constant-bank latency hides behind the MMAs of this op
                            const uint32_t d_t = op.da & 0xFFFFu;
                            uint32_t a_t = op.da >> 16;
                            uint32_t b_lo = ((base16 + op.b16) & 0x3FFFu) | (8u << 16);      // LBO = 128 B
                            const uint32_t b_hi = (op.sbo_nk & 0xFFFFu) | (1u << 14);        // SBO, descriptor version 1
                            const uint32_t idesc = op.idesc & ~1u;
                            const int nk = (int)(op.sbo_nk >> 16);
                            uint32_t acc = op.idesc & 1u;
                            for (int ks = 0; ks < nk; ++ks) {
                                mma_ts(d_t, a_t, ((uint64_t)b_hi << 32) | b_lo, idesc, acc);
                                b_lo += 16;   // two core matrices (256 B) along K
                                a_t += 8;
                                acc = 1u;
                            }
                        }
                        HINT_T2_T(t_issue);
                        if (fl & T2_LAST_IN_CHUNK) commit(bars + T2B_EMPTY + slot);
                        if (fl & T2_LAST_IN_JOB) commit(bars + T2B_MMA + job);
                    }
                }
                // the last stage's coupling must be complete before the next tile's MMAs may touch the x columns:
                // covered by bar_tile of the next tile (the epilogue warps arrive on it after storing this tile)
            }
            if (dbg && blockIdx.x == 0 && me == 0) { dbg[0] = t_tile; dbg[1] = t_prev; dbg[2] = t_epi; dbg[3] = t_chunk; dbg[4] = t_issue; dbg[6] = it; }
        }
    } else {
        // ================= epilogue warps =================
        const int wg = tid >> 7;            // 0 or 1: which half of every job's columns
        const int row = tid & 127;          // TMEM lane == sample within the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float* stg_out = reinterpret_cast<float*>(smem + P.smem_out);
        float* jx = stg_out + 128 * P.d;
        const int d = P.d, dc = P.dc;
        const float alpha = P.alpha;
        const bool round_acts = P.round_acts != 0;
        uint32_t it = 0, sctr = 0;
        long long e_xwait = 0, e_load = 0, e_waitmma = 0, e_hid = 0, e_waitfin = 0, e_fin = 0, e_store = 0, tmark = clock64();
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const long long row0 = tile * 128;
            const int rows = (int)((B - row0) < 128 ? (B - row0) : 128);
            const uint32_t buf = it & 1;
            float* stg = reinterpret_cast<float*>(smem + P.smem_in + buf * P.smem_in_bytes);
            float* stg_c = stg + 128 * d;
            mbar_wait(bars + T2B_XFULL + buf, (it >> 1) & 1);
            HINT_T2_T(e_xwait);
            if (rows < 128) {   // ragged tile: stage it by hand (zero rows beyond the batch)
                const int nval = rows * d, n = 128 * d;
                const float* g = x + row0 * d;
                for (int i = tid; i < n; i += kT2EpiThreads) stg[i] = (i < nval) ? g[i] : 0.f;
                if (dc) {
                    const int nvc = rows * dc, nc = 128 * dc;
                    const float* gc = c + row0 * dc;
                    for (int i = tid; i < nc; i += kT2EpiThreads) stg_c[i] = (i < nvc) ? gc[i] : 0.f;
                }
                named_bar_sync(1, kT2EpiThreads);
            }
            // ---- rows -> TMEM x columns (physical layout, padding = 0) ----
            for (int pc0 = 8 * wg; pc0 < P.xr; pc0 += 16) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pc = pc0 + j;
                    float val = 0.f;
                    if (pc < P.xw) { const int lc = s_xlog[pc]; if (lc >= 0) val = stg[row * d + lc]; }
                    else if (pc - P.xc < dc) val = stg_c[row * dc + (pc - P.xc)];
                    v[j] = val;
                }
                st8(lane_base + pc0, v);
            }
            wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bars + T2B_XFREE + buf); mbar_arrive(bars + T2B_TILE); }
            HINT_T2_T(e_load);
            float jacc = 0.f;
            for (int si = 0; si < nstages; ++si, ++sctr) {
                const T2Stage& S = s_stages[kRev ? si : nstages - 1 - si];
                const uint32_t par = sctr & 1;
                // hidden-layer epilogues: v = relu(v + b) in place (becomes the next layer's A operand)
                for (int j = 0; j < 3; ++j) {
                    const int ncols = S.hid[j].ncols;
                    mbar_wait(bars + T2B_MMA + j, par);
                    HINT_T2_T(e_waitmma);
                    if (ncols) {
                        fence_after_sync();
                        const float* bias = s_bias + S.hid[j].bias_off;
                        const uint32_t a0 = lane_base + S.hid[j].col0;
                        const int half = ((ncols >> 1) + 15) & ~15;
                        const int q0 = wg ? half : 0, q1 = wg ? ncols : half;
                        for (int q = q0; q < q1; q += 64) {
                            float v[4][16];
                            const int nb = (q1 - q) >> 4;   // 16-column groups left (>= 1)
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (u < nb) ld16(a0 + q + 16 * u, v[u]);
                            wait_ld();
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (u < nb) {
#pragma unroll
                                    for (int e = 0; e < 16; e += 4) {
                                        const float4 b = *reinterpret_cast<const float4*>(bias + q + 16 * u + e);
                                        v[u][e] = fmaxf(v[u][e] + b.x, 0.f); v[u][e + 1] = fmaxf(v[u][e + 1] + b.y, 0.f);
                                        v[u][e + 2] = fmaxf(v[u][e + 2] + b.z, 0.f); v[u][e + 3] = fmaxf(v[u][e + 3] + b.w, 0.f);
                                    }
                                    if (round_acts) {
#pragma unroll
                                        for (int e = 0; e < 16; ++e) v[u][e] = t2_round_tf32_pos(v[u][e]);
                                    }
                                    st16(a0 + q + 16 * u, v[u]);
                                }
                            }
                        }
                        wait_st();
                        fence_before_sync();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + T2B_EPI + j);
                    HINT_T2_T(e_hid);
                }
                // coupling epilogue (hint.py:79-84) on the lower-half x columns
                mbar_wait(bars + T2B_MMA + TC_J3S, par);
                mbar_wait(bars + T2B_MMA + TC_J3T, par);
                fence_after_sync();
                HINT_T2_T(e_waitfin);
                for (int fi = S.fin_begin + wg; fi < S.fin_end; fi += 4) {
                    float sv[2][4], tv[2][4], xv[2][4];
                    const bool two = fi + 2 < S.fin_end;
                    const T2Fin f0 = s_fins[fi];
                    const T2Fin f1 = s_fins[two ? fi + 2 : fi];
                    ld4(lane_base + f0.s_col, sv[0]); ld4(lane_base + f0.t_col, tv[0]); ld4(lane_base + f0.x_col, xv[0]);
                    if (two) { ld4(lane_base + f1.s_col, sv[1]); ld4(lane_base + f1.t_col, tv[1]); ld4(lane_base + f1.x_col, xv[1]); }
                    wait_ld();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u == 0 || two) {
                            const T2Fin& f = u ? f1 : f0;
                            const float4 bs = *reinterpret_cast<const float4*>(s_bias + f.bs_off);
                            const float4 bt = *reinterpret_cast<const float4*>(s_bias + f.bt_off);
                            const float bsv[4] = {bs.x, bs.y, bs.z, bs.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float la = alpha * t2_atan(sv[u][e] + bsv[e]);
                                const float tt = tv[u][e] + btv[e];
                                if (!kRev) { xv[u][e] = fmaf(t2_exp(la), xv[u][e], tt); jacc += la; }
                                else { xv[u][e] = (xv[u][e] - tt) * t2_exp(-la); jacc -= la; }
                            }
                            st4(lane_base + f.x_col, xv[u]);
                        }
                    }
                }
                wait_st();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + T2B_EPI + 3);
                HINT_T2_T(e_fin);
            }
            // ---- TMEM x columns -> staging rows -> global (coalesced) ----
            named_bar_sync(1, kT2EpiThreads);   // the other half's coupling writes to this row's columns; stg_out is free
            fence_after_sync();
            for (int pc0 = 8 * wg; pc0 < P.xw; pc0 += 16) {
                float v[8];
                ld8(lane_base + pc0, v);
                wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pc = pc0 + j;
                    if (pc < P.xw) { const int lc = s_xlog[pc]; if (lc >= 0) stg_out[row * d + lc] = v[j]; }
                }
            }
            if (wg == 1) jx[row] = jacc;
            fence_before_sync();
            named_bar_sync(1, kT2EpiThreads);
            {
                const int nval = rows * d;
                float* g = z + row0 * d;
                for (int i = tid * 4; i < nval; i += kT2EpiThreads * 4) {
                    if (i + 3 < nval) *reinterpret_cast<float4*>(g + i) = *reinterpret_cast<const float4*>(stg_out + i);
                    else for (int e = 0; e < 4; ++e) if (i + e < nval) g[i + e] = stg_out[i + e];
                }
                if (wg == 0 && row < rows) logdet[row0 + row] = jacc + jx[row];
            }
            HINT_T2_T(e_store);
        }
        if (dbg && blockIdx.x == 0 && tid == 0) { dbg[8] = e_xwait; dbg[9] = e_load; dbg[10] = e_waitmma; dbg[11] = e_hid; dbg[12] = e_waitfin; dbg[13] = e_fin; dbg[14] = e_store; dbg[15] = it; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(0, 512);
    }
}

// packed[i] = params[src[i]] (0 for padding); MMA operands (i < n_round) are rounded to tf32 (round-to-nearest)
__global__ void hint_pack_tc_kernel(const int* __restrict__ src, const float* __restrict__ params, float* __restrict__ packed,
                                    long long n, long long n_round) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = src[i];
        float v = s < 0 ? 0.f : params[s];
        if (i < n_round) v = tc::to_tf32(v);
        packed[i] = v;
    }
}

}  // namespace hint
