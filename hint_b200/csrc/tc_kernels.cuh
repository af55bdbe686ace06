// tcgen05 (TF32) fused-tree forward / inverse kernel: an interpreter of the static program of plan_tc.h.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles of 128 samples):
//   warps 0-7  epilogue: TMEM -> registers -> TMEM (bias + ReLU in place), the coupling epilogue
//              (atan soft clamp, exp, affine, log-det) and the tile load / store.  Warp w owns TMEM
//              lanes 32*(w%4)..+31; warps 0-3 and 4-7 split the columns of every job.
//   warps 8-11 MMA issuers: one elected thread each; the ops of every job are split between them because a
//              single thread sustains only ~1 tcgen05.mma per 80 cycles (measured: tests/cuda/umma_bench2.cu),
//              4 warps ~1 per 27 cycles.  A from TMEM, B from the ring.
//   warp 12    weight producer: one elected thread streams the packed weight images from L2 into a
//              shared-memory ring with 1-D bulk copies (mbarrier complete_tx).
// All hand-offs are mbarriers: ring full/empty, mma_done[stage][job] (tcgen05.commit), epi_done[stage][job].
#pragma once
#include "plan_tc.h"
#include "tcgen05.cuh"

namespace hint {

constexpr int kTcThreads = 256 + 32 * kTcIssuers + 32;   // 8 epilogue warps, kTcIssuers MMA warps, 1 producer warp
constexpr int kTcEpiThreads = 256;

struct TcDev {
    const TcStage* stages;
    const TcOp* ops;
    const TcChunk* chunks;
    const TcFinal* fins;
    const int* xlog;
    int nstages, nops, nchunks, nfins;
    int d, dc, xw, xc, xr;
    int slot_bytes, n_slots;
    int smem_stage_in, smem_stage_bytes, smem_tables, smem_bars, smem_ring;
    float alpha;
    int round_acts;   // round activations to tf32 (rna) in the epilogue instead of letting the MMA truncate
    int bias_base;    // first bias float in the packed buffer (== number of weight floats)
    int n_bias;
    long long* dbg;   // optional cycle breakdown of CTA 0 (nullptr in production)
};

#define HINT_TC_T(var) do { if (T.dbg) { long long now_ = clock64(); (var) += now_ - tmark; tmark = now_; } } while (0)

template <bool kRev>
__global__ void __launch_bounds__(kTcThreads, 1)
hint_fwd_tf32_kernel(TcDev T, const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ W,
                     float* __restrict__ z, float* __restrict__ logdet, long long B) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(smem + T.smem_bars + 2040);   // last word of the barrier block
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- tables -> shared memory ----
    unsigned char* tb = smem + T.smem_tables;
    TcStage* s_stages = reinterpret_cast<TcStage*>(tb);
    TcOp* s_ops = reinterpret_cast<TcOp*>(s_stages + T.nstages);
    TcChunk* s_chunks = reinterpret_cast<TcChunk*>(s_ops + T.nops);
    TcFinal* s_fins = reinterpret_cast<TcFinal*>(s_chunks + T.nchunks);
    int* s_xlog = reinterpret_cast<int*>(s_fins + T.nfins);
    float* s_bias = reinterpret_cast<float*>(s_xlog + ((T.xw + 3) & ~3));   // all biases of the block (16-byte aligned)
    {
        const int n1 = T.nstages * (int)(sizeof(TcStage) / 4), n2 = T.nops * (int)(sizeof(TcOp) / 4),
                  n3 = T.nchunks * (int)(sizeof(TcChunk) / 4), n4 = T.nfins * (int)(sizeof(TcFinal) / 4);
        for (int i = tid; i < n1; i += kTcThreads) reinterpret_cast<int*>(s_stages)[i] = reinterpret_cast<const int*>(T.stages)[i];
        for (int i = tid; i < n2; i += kTcThreads) reinterpret_cast<int*>(s_ops)[i] = reinterpret_cast<const int*>(T.ops)[i];
        for (int i = tid; i < n3; i += kTcThreads) reinterpret_cast<int*>(s_chunks)[i] = reinterpret_cast<const int*>(T.chunks)[i];
        for (int i = tid; i < n4; i += kTcThreads) reinterpret_cast<int*>(s_fins)[i] = reinterpret_cast<const int*>(T.fins)[i];
        for (int i = tid; i < T.xw; i += kTcThreads) s_xlog[i] = T.xlog[i];
        for (int i = tid; i < T.n_bias; i += kTcThreads) s_bias[i] = W[T.bias_base + i];
    }
    // ---- barriers ----
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T.smem_bars);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + T.n_slots;
    uint64_t* bar_tile = bars + 2 * T.n_slots;
    uint64_t* bar_mma = bar_tile + 1;                     // [stage][5]
    uint64_t* bar_epi = bar_mma + T.nstages * TC_NJOBS;   // [stage][4]  (0..2 hidden, 3 final)
    if (tid == 0) {
        for (int i = 0; i < T.n_slots; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, kTcIssuers); }
        mbar_init(bar_tile, 8);
        for (int i = 0; i < T.nstages * TC_NJOBS; ++i) mbar_init(bar_mma + i, kTcIssuers);
        for (int i = 0; i < T.nstages * 4; ++i) mbar_init(bar_epi + i, 8);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(&tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const long long ntiles = (B + 127) / 128;
    unsigned char* ring = smem + T.smem_ring;
    const uint32_t ring_u32 = smem_u32(ring);

    if (warp == 8 + kTcIssuers) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int si = 0; si < T.nstages; ++si) {
                    const TcStage& st = s_stages[kRev ? si : T.nstages - 1 - si];
                    for (int ch = st.chunk_begin; ch < st.chunk_end; ++ch) {
                        const uint32_t slot = cnt % T.n_slots, ph = (cnt / T.n_slots) & 1;
                        mbar_wait(bar_empty + slot, ph ^ 1);
                        const TcChunk ck = s_chunks[ch];
                        mbar_arrive_expect_tx(bar_full + slot, (uint32_t)ck.bytes);
                        bulk_g2s(ring + (size_t)slot * T.slot_bytes, W + ck.g_off, (uint32_t)ck.bytes, bar_full + slot);
                        ++cnt;
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ================= MMA issuers =================
        if (lane == 0) {
            const int me = warp - 8;
            uint32_t cnt = 0, it = 0;
            long long t_tile = 0, t_prev = 0, t_epi = 0, t_chunk = 0, t_issue = 0, t_walk = 0, tmark = clock64();
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const uint32_t tp = it & 1;
                mbar_wait(bar_tile, tp);
                fence_after_sync();
                HINT_TC_T(t_tile);
                int prev = -1;
                for (int si = 0; si < T.nstages; ++si) {
                    const int s = kRev ? si : T.nstages - 1 - si;
                    const TcStage& st = s_stages[s];
                    if (prev >= 0) {  // x columns written by the previous stage's coupling epilogue
                        mbar_wait(bar_epi + prev * 4 + 3, tp);
                        fence_after_sync();
                        HINT_TC_T(t_prev);
                    }
                    uint32_t slot = 0;
                    bool chunk_ready = false;
                    for (int oi = st.op_begin; oi < st.op_end; ++oi) {
                        const TcOp op = s_ops[oi];
                        if (op.flags & TC_FIRST_IN_CHUNK) { slot = cnt % T.n_slots; chunk_ready = false; }
                        HINT_TC_T(t_walk);
                        if (op.wait_epi >= 0) {
                            mbar_wait(bar_epi + s * 4 + op.wait_epi, tp);
                            fence_after_sync();
                            HINT_TC_T(t_epi);
                        }
                        if (op.issuer == me) {
                        if (!chunk_ready) { mbar_wait(bar_full + slot, (cnt / T.n_slots) & 1); chunk_ready = true; HINT_TC_T(t_chunk); }
                        // B descriptor: lo = (addr >> 4) | (LBO=128 >> 4) << 16 ; hi = (SBO = nk*256 >> 4) | version 1 (bit 46)
                        uint32_t b_lo = ((ring_u32 + slot * (uint32_t)T.slot_bytes + (uint32_t)op.b_off * 4) >> 4) | (8u << 16);
                        const uint32_t b_hi = ((uint32_t)op.nk * 16) | (1u << 14);
                        uint32_t a_t = tbase + op.a_col;
                        const uint32_t d_t = tbase + op.d_col;
                        uint32_t acc = (op.flags & TC_ACCUM) ? 1u : 0u;
                        for (int ks = 0; ks < op.nk; ++ks) {
                            mma_ts(d_t, a_t, ((uint64_t)b_hi << 32) | b_lo, op.idesc, acc);
                            b_lo += 16;   // two core matrices (256 B) along K
                            a_t += 8;
                            acc = 1u;
                        }
                        HINT_TC_T(t_issue);
                        }
                        if (op.commit_job >= 0) commit(bar_mma + s * TC_NJOBS + op.commit_job);
                        if (op.flags & TC_LAST_IN_CHUNK) { commit(bar_empty + slot); ++cnt; }
                    }
                    prev = s;
                }
            }
            if (T.dbg && blockIdx.x == 0 && me == 0) {
                T.dbg[0] = t_tile; T.dbg[1] = t_prev; T.dbg[2] = t_epi; T.dbg[3] = t_chunk; T.dbg[4] = t_issue; T.dbg[5] = t_walk; T.dbg[6] = it;
            }
        }
    } else {
        // ================= epilogue warps =================
        const int wg = tid >> 7;            // 0 or 1: which half of the column groups
        const int row = tid & 127;          // TMEM lane == sample within the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float* stg = reinterpret_cast<float*>(smem + T.smem_stage_in);
        float* stg_c = stg + 128 * T.d;
        float* jx = reinterpret_cast<float*>(smem + T.smem_stage_in + T.smem_stage_bytes);  // 128 floats of scratch
        uint32_t it = 0;
        long long e_load = 0, e_waitmma = 0, e_hid = 0, e_waitfin = 0, e_fin = 0, e_store = 0, tmark = clock64();
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t tp = it & 1;
            const long long row0 = tile * 128;
            const int rows = (int)((B - row0) < 128 ? (B - row0) : 128);
            // ---- stage the tile (coalesced) ----
            {
                const int nval = rows * T.d, n = 128 * T.d;
                const float* g = x + row0 * T.d;
                for (int i = tid * 4; i < n; i += kTcEpiThreads * 4) {
                    if (i + 3 < nval) *reinterpret_cast<float4*>(stg + i) = __ldg(reinterpret_cast<const float4*>(g + i));
                    else for (int e = 0; e < 4; ++e) stg[i + e] = (i + e < nval) ? g[i + e] : 0.f;
                }
                if (T.dc) {
                    const int nvc = rows * T.dc, nc = 128 * T.dc;
                    const float* gc = c + row0 * T.dc;
                    for (int i = tid; i < nc; i += kTcEpiThreads) stg_c[i] = (i < nvc) ? gc[i] : 0.f;
                }
            }
            named_bar_sync(1, kTcEpiThreads);
            // ---- rows -> TMEM x columns (physical layout, padding = 0) ----
            for (int pc0 = 8 * wg; pc0 < T.xr; pc0 += 16) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pc = pc0 + j;
                    float val = 0.f;
                    if (pc < T.xw) { const int lc = s_xlog[pc]; if (lc >= 0) val = stg[row * T.d + lc]; }
                    else if (pc - T.xc < T.dc) val = stg_c[row * T.dc + (pc - T.xc)];
                    v[j] = val;
                }
                st8(tbase + lane_base + pc0, v);
            }
            wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tile);
            HINT_TC_T(e_load);
            float jacc = 0.f;
            for (int si = 0; si < T.nstages; ++si) {
                const int s = kRev ? si : T.nstages - 1 - si;
                const TcStage& st = s_stages[s];
                // hidden-layer epilogues: v = relu(v + b), in place (becomes the next layer's A operand)
                for (int j = 0; j < 3; ++j) {
                    const TcHidden h = st.hid[j];
                    if (h.ncols == 0) continue;
                    mbar_wait(bar_mma + s * TC_NJOBS + j, tp);
                    fence_after_sync();
                    HINT_TC_T(e_waitmma);
                    const float* bias = s_bias + (h.bias_off - T.bias_base);
                    for (int q = 16 * wg; q < h.ncols; q += 32) {
                        float v[16];
                        const uint32_t a = tbase + lane_base + h.col0 + q;
                        ld16(a, v);
                        float bb[16];
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + q + e);
                            bb[e] = b.x; bb[e + 1] = b.y; bb[e + 2] = b.z; bb[e + 3] = b.w;
                        }
                        wait_ld();
                        if (T.round_acts) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) v[e] = to_tf32(fmaxf(v[e] + bb[e], 0.f));
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e] + bb[e], 0.f);
                        }
                        st16(a, v);
                    }
                    wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_epi + s * 4 + j);
                    HINT_TC_T(e_hid);
                }
                // coupling epilogue (hint.py:79-84) on the lower-half x columns
                if (st.has_job[TC_J3S]) mbar_wait(bar_mma + s * TC_NJOBS + TC_J3S, tp);
                if (st.has_job[TC_J3T]) mbar_wait(bar_mma + s * TC_NJOBS + TC_J3T, tp);
                fence_after_sync();
                HINT_TC_T(e_waitfin);
                for (int fi = st.fin_begin + wg; fi < st.fin_end; fi += 2) {
                    const TcFinal f = s_fins[fi];
                    float sv[4], tv[4], xv[4];
                    ld4(tbase + lane_base + f.s_col, sv);
                    ld4(tbase + lane_base + f.t_col, tv);
                    ld4(tbase + lane_base + f.x_col, xv);
                    const float4 bs = *reinterpret_cast<const float4*>(s_bias + (f.bs_off - T.bias_base));
                    const float4 bt = *reinterpret_cast<const float4*>(s_bias + (f.bt_off - T.bias_base));
                    wait_ld();
                    const float bsv[4] = {bs.x, bs.y, bs.z, bs.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float la = T.alpha * atanf(sv[e] + bsv[e]);
                        const float tt = tv[e] + btv[e];
                        if (!kRev) { xv[e] = expf(la) * xv[e] + tt; jacc += la; }
                        else { xv[e] = (xv[e] - tt) / expf(la); jacc -= la; }
                    }
                    st4(tbase + lane_base + f.x_col, xv);
                }
                wait_st();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_epi + s * 4 + 3);
                HINT_TC_T(e_fin);
            }
            // ---- TMEM x columns -> staging rows -> global (coalesced) ----
            named_bar_sync(1, kTcEpiThreads);   // the other half's coupling writes to this row's columns
            fence_after_sync();
            for (int pc0 = 8 * wg; pc0 < T.xw; pc0 += 16) {
                float v[8];
                ld8(tbase + lane_base + pc0, v);
                wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pc = pc0 + j;
                    if (pc < T.xw) { const int lc = s_xlog[pc]; if (lc >= 0) stg[row * T.d + lc] = v[j]; }
                }
            }
            if (wg == 1) jx[row] = jacc;
            fence_before_sync();
            named_bar_sync(1, kTcEpiThreads);
            {
                const int nval = rows * T.d;
                float* g = z + row0 * T.d;
                for (int i = tid * 4; i < nval; i += kTcEpiThreads * 4) {
                    if (i + 3 < nval) *reinterpret_cast<float4*>(g + i) = *reinterpret_cast<const float4*>(stg + i);
                    else for (int e = 0; e < 4; ++e) if (i + e < nval) g[i + e] = stg[i + e];
                }
                if (wg == 0 && row < rows) logdet[row0 + row] = jacc + jx[row];
            }
            named_bar_sync(1, kTcEpiThreads);
            HINT_TC_T(e_store);
        }
        if (T.dbg && blockIdx.x == 0 && tid == 0) {
            T.dbg[8] = e_load; T.dbg[9] = e_waitmma; T.dbg[10] = e_hid; T.dbg[11] = e_waitfin; T.dbg[12] = e_fin; T.dbg[13] = e_store; T.dbg[14] = it;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tbase, 512);
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
