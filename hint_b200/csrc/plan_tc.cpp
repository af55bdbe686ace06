// Generator of the static tcgen05 (TF32) program of one HINT block — see plan_tc.h.  Host only.
#include "plan_tc.h"

#include <algorithm>
#include <cstring>
#include <functional>

namespace hint {

namespace {

inline int r16(int v) { return (v + 15) & ~15; }
inline int canon_off(int r, int k, int kpad) { return (r >> 3) * (kpad * 8) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

struct NodeLay {
    int hp8 = 0;
    bool bd = false;          // s and t fused block-diagonally in layers 2 and 3 (h <= 8 and <= 8 outputs)
    int w1 = 0, w2a = 0, w2b = 0, w3 = 0;
    int o1 = 0, o2a = 0, o2b = 0, o3 = 0;   // TMEM columns (absolute) once placed in a stage
    int plo = 0, pw = 0;      // physical range of the node in the x columns
    int ulo = 0, uw = 0;      // physical range read by layer 1 (the upper half)
    int po = 0, pwo = 0;      // physical range written by the coupling (the lower half, 4-aligned start)
    int s3 = 0, t3 = 0;       // offsets of the s / t outputs inside the node's layer-3 region
};

}  // namespace

void build_tc_schedule(const Plan& p, TcSchedule& t) {
    t = TcSchedule();
    t.d = p.d;
    t.dc = p.dc;
    const int d = p.d, dc = p.dc;
    const int nn = (int)p.nodes.size();
    auto fail = [&](const std::string& why) { t.ok = false; t.why = why; };
    auto P = [&](int node, int net, int layer, int kind) { return p.param_offsets[(size_t)node * 12 + net * 6 + layer * 2 + kind]; };

    // ---- physical x layout: every leaf padded to a multiple of 4 columns, so all node ranges are 4-aligned ----
    t.xphys.assign(d, -1);
    std::vector<NodeLay> lay(nn);
    int col = 0;
    for (int i = 0; i < nn; ++i) {
        const auto& n = p.nodes[i];
        if (!n.leaf) continue;
        for (int c = n.lo; c < n.hi; ++c) t.xphys[c] = col + (c - n.lo);
        col += round4(n.hi - n.lo);
    }
    t.xw = col;
    t.xlog.assign(col, -1);
    for (int c = 0; c < d; ++c) t.xlog[t.xphys[c]] = c;
    const int dcp8 = round8(dc);
    t.xc = t.xw;
    t.xr = r16(t.xw + std::max(dcp8, 8));   // >= 8 finite (zero / condition) columns behind the x columns
    for (int i = nn - 1; i >= 0; --i) {  // children have larger pre-order indices
        const auto& n = p.nodes[i];
        NodeLay& L = lay[i];
        L.plo = t.xphys[n.lo];
        L.pw = n.leaf ? round4(n.hi - n.lo) : lay[n.upper].pw + lay[n.lower].pw;
        if (n.k == 0) return fail("node with an empty upper half");
        L.ulo = L.plo;
        L.uw = n.leaf ? n.k : lay[n.upper].pw;
        L.po = n.leaf ? ((L.plo + n.k) & ~3) : lay[n.lower].plo;
        L.pwo = L.plo + L.pw - L.po;
        L.hp8 = round8(n.h);
        L.bd = (L.hp8 == 8 && L.pwo <= 8);
        L.w1 = r16(2 * L.hp8);
        if (L.bd) { L.w2a = 16; L.w2b = 0; L.w3 = 16; L.s3 = 0; L.t3 = 8; }
        else { L.w2a = L.w2b = r16(n.h); L.w3 = 2 * r16(L.pwo); L.s3 = 0; L.t3 = r16(L.pwo); }
    }
    if (t.xr + 64 > t.tmem_cols) return fail("input width too large for the TMEM-resident x tile");

    // ---- stages: greedy grouping of each level's nodes under the TMEM column budget ----
    const int cap = t.tmem_cols - t.xr;
    std::vector<std::vector<int>> groups;
    for (int depth = 0; depth <= p.max_depth; ++depth) {
        std::vector<int> cur;
        int used = 0;
        for (int i = 0; i < nn; ++i) {
            if (p.nodes[i].depth != depth) continue;
            const int w = lay[i].w1 + lay[i].w2a + lay[i].w2b + lay[i].w3;
            if (w > cap) return fail("hidden width too large for the TMEM-resident activations of the TF32 kernel");
            if (!cur.empty() && used + w > cap) { groups.push_back(cur); cur.clear(); used = 0; }
            cur.push_back(i);
            used += w;
        }
        if (!cur.empty()) groups.push_back(cur);
    }

    // ---- shared-memory budget: staging (2 x tile of x and c), weight ring, tables, barriers ----
    const int stage_bytes = ((128 * (d + dc) * 4) + 127) & ~127;
    const int avail = kSmemMax - 2 * stage_bytes - 40 * 1024;
    int slot = std::min(64 * 1024, avail / 3) & ~1023;
    if (slot < 16 * 1024) return fail("not enough shared memory for the weight ring");
    t.slot_bytes = slot;

    std::vector<int32_t> wsrc;   // weight images
    std::vector<int32_t> bsrc;   // biases (appended after the weights)
    auto alloc_bias = [&](int n) { int o = (int)bsrc.size(); bsrc.resize(bsrc.size() + ((n + 3) & ~3), -1); return o; };

    for (const auto& g : groups) {
        TcStage st{};
        st.op_begin = (int)t.ops.size();
        st.chunk_begin = (int)t.chunks.size();
        st.fin_begin = (int)t.fins.size();
        // place regions
        int w1 = 0, w2a = 0, w2b = 0, w3 = 0;
        for (int i : g) { w1 += lay[i].w1; w2a += lay[i].w2a; w2b += lay[i].w2b; w3 += lay[i].w3; }
        const int R1 = t.xr, R2a = R1 + w1, R2b = R2a + w2a, R3 = R2b + w2b;
        int a1 = R1, a2a = R2a, a2b = R2b, a3 = R3;
        for (int i : g) {
            lay[i].o1 = a1; a1 += lay[i].w1;
            lay[i].o2a = a2a; a2a += lay[i].w2a;
            lay[i].o2b = a2b; a2b += lay[i].w2b;
            lay[i].o3 = a3; a3 += lay[i].w3;
        }
        st.hid[0] = TcHidden{R1, w1, alloc_bias(w1), 0};
        st.hid[1] = TcHidden{R2a, w2a, alloc_bias(w2a), 0};
        st.hid[2] = TcHidden{R2b, w2b, w2b ? alloc_bias(w2b) : 0, 0};

        long long chunk_g = (long long)wsrc.size();
        int chunk_bytes = 0;
        bool chunk_open = false;
        auto close_chunk = [&]() {
            if (!chunk_open) return;
            t.chunks.push_back(TcChunk{chunk_g, chunk_bytes, 0});
            t.ops.back().flags |= TC_LAST_IN_CHUNK;
            chunk_open = false;
        };
        int job_first_op[TC_NJOBS];
        for (int& v : job_first_op) v = -1;
        int job_last_op[TC_NJOBS];
        for (int& v : job_last_op) v = -1;
        bool bad = false;
        int issuer_load[TC_NJOBS][kTcIssuers] = {};
        // emit D[:, d_col .. d_col+n_rows) (+)= A[:, a_col .. a_col+8nk) * B^T with B(n,k) = params[fill(n,k)] (or 0)
        auto emit = [&](int job, int d_col, int a_col, int n_rows, int nk, bool accum, const std::function<long long(int, int)>& fill) {
            if (nk == 0) return;
            const int kpad = 8 * nk;
            int max_rows = (slot / (kpad * 4)) & ~15;
            if (max_rows < 16) { bad = true; return; }
            for (int r0 = 0; r0 < n_rows; r0 += max_rows) {
                const int rows = std::min(max_rows, n_rows - r0);
                const int img = rows * kpad * 4;
                TcOp op{};
                if (!chunk_open || chunk_bytes + img > slot) {
                    close_chunk();
                    chunk_g = (long long)wsrc.size();
                    chunk_bytes = 0;
                    chunk_open = true;
                    op.flags |= TC_FIRST_IN_CHUNK;
                }
                op.d_col = d_col + r0;
                op.a_col = a_col;
                op.b_off = chunk_bytes / 4;
                op.nk = nk;
                op.n_rows = rows;
                op.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(rows >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
                if (accum) op.flags |= TC_ACCUM;
                op.wait_epi = -1;
                op.job = (signed char)job;
                op.commit_job = -1;
                if (accum && !t.ops.empty()) {
                    op.issuer = t.ops.back().issuer;       // accumulates onto the previous op's D: same thread, program order
                } else {
                    int best = 0;
                    for (int q = 1; q < kTcIssuers; ++q)
                        if (issuer_load[job][q] < issuer_load[job][best]) best = q;
                    op.issuer = (unsigned char)best;
                }
                issuer_load[job][op.issuer] += nk + 2;
                const size_t base = wsrc.size();
                wsrc.resize(base + (size_t)rows * kpad, -1);
                for (int r = 0; r < rows; ++r)
                    for (int k = 0; k < kpad; ++k) {
                        const long long src = fill(r0 + r, k);
                        if (src >= 0) wsrc[base + canon_off(r, k, kpad)] = (int32_t)src;
                    }
                chunk_bytes += img;
                if (job_first_op[job] < 0) job_first_op[job] = (int)t.ops.size();
                job_last_op[job] = (int)t.ops.size();
                t.ops.push_back(op);
            }
        };

        // ---- J1: layer 1, s and t fused on the output side (shared A = x_upper [, c]) ----
        for (int i : g) {
            const auto& n = p.nodes[i];
            const NodeLay& L = lay[i];
            auto unit = [&](int r, int& net, int& u) { net = r < L.hp8 ? 0 : 1; u = r - net * L.hp8; return r < 2 * L.hp8 && u < n.h; };
            emit(TC_J1, L.o1, L.ulo, L.w1, round8(L.uw) / 8, false, [&](int r, int k) -> long long {
                int net, u;
                if (!unit(r, net, u)) return -1;
                if (L.ulo + k >= t.xw) return -1;
                const int c = t.xlog[L.ulo + k];
                if (c < n.lo || c >= n.lo + n.k) return -1;
                return P(i, net, 0, 0) + (long long)u * n.cin + (c - n.lo);
            });
            if (dc > 0)
                emit(TC_J1, L.o1, t.xc, L.w1, dcp8 / 8, true, [&](int r, int k) -> long long {
                    int net, u;
                    if (!unit(r, net, u) || k >= dc) return -1;
                    return P(i, net, 0, 0) + (long long)u * n.cin + n.k + k;
                });
            for (int r = 0; r < L.w1; ++r) {
                int net, u;
                if (unit(r, net, u)) bsrc[st.hid[0].bias_off + (L.o1 - R1) + r] = (int32_t)(P(i, net, 0, 1) + u);
            }
        }
        // ---- J2S / J2T: layer 2 ----
        for (int pass = 0; pass < 2; ++pass)
            for (int i : g) {
                const auto& n = p.nodes[i];
                const NodeLay& L = lay[i];
                if (L.bd) {
                    if (pass == 1) continue;
                    emit(TC_J2S, L.o2a, L.o1, 16, 2, false, [&](int r, int k) -> long long {
                        const int net = r / 8, u = r % 8, kn = k / 8, ku = k % 8;
                        if (net != kn || u >= n.h || ku >= n.h) return -1;
                        return P(i, net, 1, 0) + (long long)u * n.h + ku;
                    });
                    for (int r = 0; r < 16; ++r)
                        if (r % 8 < n.h) bsrc[st.hid[1].bias_off + (L.o2a - R2a) + r] = (int32_t)(P(i, r / 8, 1, 1) + r % 8);
                } else {
                    const int net = pass;
                    emit(net == 0 ? TC_J2S : TC_J2T, net == 0 ? L.o2a : L.o2b, L.o1 + net * L.hp8, L.w2a, L.hp8 / 8, false,
                         [&](int r, int k) -> long long {
                             if (r >= n.h || k >= n.h) return -1;
                             return P(i, net, 1, 0) + (long long)r * n.h + k;
                         });
                    const int boff = net == 0 ? st.hid[1].bias_off + (L.o2a - R2a) : st.hid[2].bias_off + (L.o2b - R2b);
                    for (int r = 0; r < n.h; ++r) bsrc[boff + r] = (int32_t)(P(i, net, 1, 1) + r);
                }
            }
        // ---- J3S / J3T: layer 3; output row r <-> physical x column po + r (rows outside the lower half are zero) ----
        for (int pass = 0; pass < 2; ++pass)
            for (int i : g) {
                const auto& n = p.nodes[i];
                const NodeLay& L = lay[i];
                auto out_row = [&](int r) {  // -> index into the subnet output (0..cout-1) or -1
                    if (r >= L.pwo) return -1;
                    const int c = t.xlog[L.po + r];
                    return (c >= n.lo + n.k && c < n.hi) ? c - (n.lo + n.k) : -1;
                };
                if (L.bd) {
                    if (pass == 1) continue;
                    emit(TC_J3S, L.o3, L.o2a, 16, 2, false, [&](int r, int k) -> long long {
                        const int net = r / 8, kn = k / 8, ku = k % 8;
                        const int o = out_row(r % 8);
                        if (net != kn || o < 0 || ku >= n.h) return -1;
                        return P(i, net, 2, 0) + (long long)o * n.h + ku;
                    });
                } else {
                    const int net = pass;
                    emit(net == 0 ? TC_J3S : TC_J3T, L.o3 + (net == 0 ? L.s3 : L.t3), net == 0 ? L.o2a : L.o2b, r16(L.pwo), L.hp8 / 8, false,
                         [&](int r, int k) -> long long {
                             const int o = out_row(r);
                             if (o < 0 || k >= n.h) return -1;
                             return P(i, net, 2, 0) + (long long)o * n.h + k;
                         });
                }
                if (pass == 0) {
                    const int bs = alloc_bias(L.pwo), bt = alloc_bias(L.pwo);
                    for (int r = 0; r < L.pwo; ++r) {
                        const int o = out_row(r);
                        if (o >= 0) { bsrc[bs + r] = (int32_t)(P(i, 0, 2, 1) + o); bsrc[bt + r] = (int32_t)(P(i, 1, 2, 1) + o); }
                    }
                    for (int q = 0; q < L.pwo; q += 4)
                        t.fins.push_back(TcFinal{L.o3 + L.s3 + q, L.o3 + L.t3 + q, L.po + q, 4, bs + q, bt + q, 0, 0});
                }
            }
        if (bad) return fail("a weight matrix row block exceeds one ring slot");
        close_chunk();
        // dependencies and completion signals
        static const int dep[TC_NJOBS] = {-1, TC_J1, TC_J1, TC_J2S, TC_J2T};
        for (int j = 0; j < TC_NJOBS; ++j) {
            st.has_job[j] = job_first_op[j] >= 0;
            if (!st.has_job[j]) continue;
            t.ops[job_first_op[j]].wait_epi = (signed char)dep[j];
            t.ops[job_last_op[j]].commit_job = (signed char)j;
        }
        if (!st.has_job[TC_J2T]) st.hid[2].ncols = 0;
        st.op_end = (int)t.ops.size();
        st.chunk_end = (int)t.chunks.size();
        st.fin_end = (int)t.fins.size();
        t.stages.push_back(st);
    }

    // ---- finalise the packed buffer: [weight images | biases] ----
    t.n_weight_floats = (long long)wsrc.size();
    t.pack_src = wsrc;
    t.pack_src.insert(t.pack_src.end(), bsrc.begin(), bsrc.end());
    while (t.pack_src.size() % 4) t.pack_src.push_back(-1);
    if (t.n_weight_floats % 4) return fail("internal: weight images must be 16-byte multiples");
    t.n_packed = (long long)t.pack_src.size();
    const int boff = (int)t.n_weight_floats;
    for (auto& st : t.stages)
        for (auto& h : st.hid) h.bias_off += boff;
    for (auto& f : t.fins) { f.bs_off += boff; f.bt_off += boff; }

    // ---- shared-memory layout ----
    int off = 0;
    t.smem_stage_in = off; t.smem_stage_bytes = stage_bytes; off += 2 * stage_bytes;
    t.smem_tables = off;
    t.smem_tables_bytes = (int)(t.stages.size() * sizeof(TcStage) + t.ops.size() * sizeof(TcOp) + t.chunks.size() * sizeof(TcChunk) +
                                t.fins.size() * sizeof(TcFinal) + t.xw * 4 + 64 + (t.n_packed - t.n_weight_floats) * 4);
    t.smem_tables_bytes = (t.smem_tables_bytes + 127) & ~127;
    off += t.smem_tables_bytes;
    t.smem_bars = off;
    off += 2048;
    off = (off + 1023) & ~1023;
    t.smem_ring = off;
    t.n_slots = std::min(4, (kSmemMax - off) / slot);
    if (t.n_slots < 2) return fail("not enough shared memory for a double-buffered weight ring");
    off += t.n_slots * slot;
    t.smem_bytes = (size_t)off;
    if (t.stages.size() * 9 + 2 * 4 + 1 > 250) return fail("too many stages for the barrier table");
    t.ok = true;
}

// ---- v2 encoding -----------------------------------------------------------------------------------------------
void build_tc2_program(const Plan& p, const TcSchedule& t, T2Host& out) {
    out.ok = false;
    out.why.clear();
    out.smem_bytes = 0;
    if (!t.ok) { out.why = t.why; return; }
    T2Prog& P = out.prog;
    std::memset(&P, 0, sizeof(P));
    auto fail = [&](const std::string& why) { out.ok = false; out.why = why; };
    if ((int)t.stages.size() > kT2MaxStages || (int)t.ops.size() > kT2MaxOps || (int)t.fins.size() > kT2MaxFins ||
        (int)t.chunks.size() > kT2MaxChunks || t.xw > kT2MaxXw)
        return fail("program too large for the kernel-parameter space of the TF32 kernel");
    P.nstages = (int)t.stages.size();
    P.nfins = (int)t.fins.size();
    P.d = t.d; P.dc = t.dc; P.xw = t.xw; P.xc = t.xc; P.xr = t.xr;
    P.slot_bytes = t.slot_bytes;
    P.bias_base = (int)t.n_weight_floats;
    P.n_bias = (int)(t.n_packed - t.n_weight_floats);
    P.alpha = p.alpha;
    P.round_acts = 1;
    for (int i = 0; i < t.xw; ++i) P.xlog[i] = (int16_t)t.xlog[i];
    for (size_t i = 0; i < t.chunks.size(); ++i) {
        if (t.chunks[i].g_off % 4 || t.chunks[i].bytes % 16) return fail("internal: unaligned weight chunk");
        P.chunks[i] = T2Chunk{(uint32_t)(t.chunks[i].g_off / 4), (uint32_t)t.chunks[i].bytes};
    }
    for (size_t i = 0; i < t.fins.size(); ++i) {
        const TcFinal& f = t.fins[i];
        P.fins[i] = T2Fin{(uint16_t)f.s_col, (uint16_t)f.t_col, (uint16_t)f.x_col, (uint16_t)(f.bs_off - P.bias_base), (uint16_t)(f.bt_off - P.bias_base)};
    }
    if (P.n_bias > 65535) return fail("too many biases for the 16-bit offsets of the TF32 program");
    int nseg = 0, nop = 0;
    for (int s = 0; s < P.nstages; ++s) {
        const TcStage& st = t.stages[s];
        T2Stage& S = P.stages[s];
        S.seg_begin = (uint16_t)nseg;
        S.fin_begin = (uint16_t)st.fin_begin; S.fin_end = (uint16_t)st.fin_end;
        S.chunk_begin = (uint16_t)st.chunk_begin; S.chunk_end = (uint16_t)st.chunk_end;
        for (int j = 0; j < 3; ++j)
            S.hid[j] = T2Hidden{(uint16_t)st.hid[j].col0, (uint16_t)st.hid[j].ncols, (uint16_t)(st.hid[j].ncols ? st.hid[j].bias_off - P.bias_base : 0), 0};
        // runs of ops with the same (job, chunk); every job gets at least one (possibly empty) segment so that every
        // barrier of the kernel completes exactly once per stage
        int oi = st.op_begin;
        for (int job = 0; job < TC_NJOBS; ++job) {
            bool first = true;
            int last_seg = -1;
            while (oi < st.op_end && t.ops[oi].job == job) {
                int oe = oi + 1;
                while (oe < st.op_end && t.ops[oe].job == job && !(t.ops[oe].flags & TC_FIRST_IN_CHUNK)) ++oe;
                if (nseg >= kT2MaxSegs) return fail("program too large for the kernel-parameter space of the TF32 kernel");
                T2Seg& G = P.segs[nseg];
                G.job = (uint8_t)job;
                G.flags = (uint8_t)((first ? T2_FIRST_IN_JOB : 0) | ((t.ops[oi].flags & TC_FIRST_IN_CHUNK) ? T2_FIRST_IN_CHUNK : 0) |
                                    ((t.ops[oe - 1].flags & TC_LAST_IN_CHUNK) ? T2_LAST_IN_CHUNK : 0));
                for (int q = 0; q < kTcIssuers; ++q) {
                    G.op_ofs[q] = (uint16_t)nop;
                    for (int o = oi; o < oe; ++o) {
                        const TcOp& op = t.ops[o];
                        if (op.issuer != q) continue;
                        if ((op.b_off * 4) % 16) return fail("internal: unaligned weight image");
                        const uint32_t sbo = (uint32_t)op.nk * 256;   // 8 rows x (8*nk) floats
                        P.ops[nop++] = T2Op{(uint32_t)op.d_col | ((uint32_t)op.a_col << 16), (uint32_t)(op.b_off * 4) >> 4,
                                            (sbo >> 4) | ((uint32_t)op.nk << 16), op.idesc | ((op.flags & TC_ACCUM) ? 1u : 0u)};
                    }
                }
                G.op_ofs[kTcIssuers] = (uint16_t)nop;
                first = false;
                last_seg = nseg++;
                oi = oe;
            }
            if (last_seg < 0) {
                if (nseg >= kT2MaxSegs) return fail("program too large for the kernel-parameter space of the TF32 kernel");
                T2Seg& G = P.segs[nseg];
                for (int q = 0; q <= kTcIssuers; ++q) G.op_ofs[q] = (uint16_t)nop;
                G.job = (uint8_t)job;
                G.flags = T2_FIRST_IN_JOB;
                last_seg = nseg++;
            }
            P.segs[last_seg].flags |= T2_LAST_IN_JOB;
        }
        if (oi != st.op_end) return fail("internal: ops of a stage are not ordered by job");
        S.seg_end = (uint16_t)nseg;
    }
    // ---- shared memory: [barriers 1 KB][epilogue tables][biases][x/c staging in x2][z staging + log-det scratch][ring] ----
    int off = 1024;
    P.smem_tab = off;   // copies of the epilogue tables: stages | fins | xlog
    off += (int)((P.nstages * sizeof(T2Stage) + P.nfins * sizeof(T2Fin) + t.xw * sizeof(int16_t) + 127) & ~size_t(127)) + 128;
    P.smem_bias = off; off += ((P.n_bias * 4) + 127) & ~127;
    P.smem_in_bytes = ((128 * (t.d + t.dc) * 4) + 127) & ~127;
    P.smem_in = off; off += 2 * P.smem_in_bytes;
    P.smem_out = off; off += ((128 * t.d * 4 + 512) + 127) & ~127;
    off = (off + 1023) & ~1023;
    P.smem_ring = off;
    P.n_slots = std::min(6, (kSmemMax - off) / t.slot_bytes);
    if (P.n_slots < 2) return fail("not enough shared memory for a double-buffered weight ring");
    off += P.n_slots * t.slot_bytes;
    out.smem_bytes = (size_t)off;
    out.ok = true;
}

}  // namespace hint
