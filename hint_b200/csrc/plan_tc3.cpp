// Planner of the tcgen05 training kernel (see plan_tc3.h for the machine model).
//
// What each piece restates from the reference (/root/reference/hint.py):
//   groups / levels ...... the recursion of hint.py:47-52 flattened: nodes of equal depth never interact (hint.py:72-73)
//   phase order .......... inverse sweep, root coupling first (hint.py:85-88): the backward visits nodes in that order
//   subnet layers ........ Linear-ReLU-Linear-ReLU-Linear (hint.py:10-13), s and t share the input (hint.py:76-77)
//   coupling ............. hint.py:79-84 and its derivative (formulas: oracle/hint_oracle.py backward_from_output)
#include "plan_tc3.h"

#include <algorithm>
#include <cstring>

namespace hint {

namespace {

constexpr int kTmemCols = 512;
constexpr int kHpMaxMulti = 128;   // hidden columns of a multi-node group (one M tile of the weight-gradient GEMMs)
constexpr int kHpMaxSingle = 160;  // a single node may be wider (second M tile)
constexpr int kHpMaxTransport = 400; // forward / inverse programs: the first hidden layer must fit TMEM next to a chunk of the second
constexpr int kKaMax = 32, kOwMax = 32;
constexpr int kKaMaxTransport = 64, kOwMaxTransport = 64;   // no images / accumulators keyed to them in the transport programs

enum ResKind { R_TMEM = 0, R_IMG0 = 1 /* .. R_IMG0 + kT3Imgs - 1 */ };
struct Res { int kind, lo, hi; };
struct Acc { std::vector<Res> rd, wr; };

bool overlap(const Res& a, const Res& b) { return a.kind == b.kind && a.lo < b.hi && b.lo < a.hi; }
bool conflicts(const Acc& x, const Acc& y) {   // x later than y
    for (const Res& r : x.rd) for (const Res& w : y.wr) if (overlap(r, w)) return true;
    for (const Res& w : x.wr) {
        for (const Res& r : y.rd) if (overlap(w, r)) return true;
        for (const Res& w2 : y.wr) if (overlap(w, w2)) return true;
    }
    return false;
}

int64_t poffc(const Plan& p, int node, int net, int layer, int kind) {
    return p.param_offsets[(size_t)node * 12 + net * 6 + layer * 2 + kind];
}

struct BOp {                 // one B operand [N][K], K a multiple of 8, N a multiple of 8
    int N = 0, K = 0;
    std::vector<int32_t> src;   // [n*K + k] -> parameter index or -1
    void init(int n, int k) { N = n; K = k; src.assign((size_t)n * k, -1); }
    int32_t& at(int n, int k) { return src[(size_t)n * K + k]; }
};
struct ASeg { int col, nk; };   // A operand: nk K steps starting at TMEM column col

struct Builder {
    const Plan& p;
    T3Plan& t;
    std::vector<Acc> macc, eacc;        // access sets per MMA record / epilogue step
    std::vector<int> mtime, etime;      // logical timestamps
    int clock = 0;
    int open_chunk = -1;                // chunk whose records are being emitted
    // Deferred accumulator flushes.  A flush only reads its accumulator columns, so it may run any time before those columns
    // are written again: the planner parks it here and re-emits it where the epilogue warps would otherwise idle behind a long
    // MMA (the hidden-layer GEMMs of the NEXT forward chain), or - at the latest - right before a record that overwrites it.
    std::vector<std::pair<T3Epi, Acc>> pending;
    Builder(const Plan& p_, T3Plan& t_) : p(p_), t(t_) {}
    void defer(const T3Epi& e, const Acc& a) { pending.emplace_back(e, a); }
    void drain(int n) {   // emit up to n parked flushes (oldest first); n < 0: all
        while (!pending.empty() && n != 0) {
            push_epi(pending.front().first, pending.front().second);
            pending.erase(pending.begin());
            if (n > 0) --n;
        }
    }

    // ---- weights ----
    // Appends the operand as K slabs (each an independent canonical block that fits one ring slot) and emits the
    // tcgen05.mma records that multiply the A segments with it into D.
    void gemm_ts(const BOp& B, const std::vector<ASeg>& aseg, int d_col, bool commit, bool zero_first = true) {
        const int max_kc = std::max(8, (t.slot_bytes / (B.N * 4)) / 8 * 8);
        // K steps in order: (a_col of each step)
        std::vector<int> acols;
        for (const ASeg& s : aseg) for (int i = 0; i < s.nk; ++i) acols.push_back(s.col + 8 * i);
        const int nk_total = B.K / 8;
        bool first = zero_first;
        for (int k0 = 0; k0 < nk_total;) {
            const int kc = std::min(max_kc / 8, nk_total - k0);      // K steps in this slab
            const int Kc = kc * 8;
            // pack the slab
            T3Chunk ck;
            ck.g_off = (uint32_t)t.pack_src.size();
            ck.bytes = (uint32_t)(B.N * Kc * 4);
            t.pack_src.resize(t.pack_src.size() + (size_t)B.N * Kc, -1);
            for (int n = 0; n < B.N; ++n)
                for (int k = 0; k < Kc; ++k) t.pack_src[ck.g_off + t3_canon_off(n, k, Kc)] = B.src[(size_t)n * B.K + k0 * 8 + k];
            t.chunks.push_back(ck);
            // records: break at A discontinuities
            int s0 = 0;
            while (s0 < kc) {
                int s1 = s0 + 1;
                while (s1 < kc && acols[k0 + s1] == acols[k0 + s1 - 1] + 8) ++s1;
                T3Mma m{};
                m.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(B.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                m.b_off = (uint32_t)(s0 * 256);            // two core matrices (256 B) per K step
                m.d_col = (uint16_t)d_col;
                m.a_col = (uint16_t)acols[k0 + s0];
                m.nk = (uint16_t)(s1 - s0);
                m.b_sbo16 = (uint16_t)((Kc * 32) >> 4);
                m.flags = 0;
                if (first) m.flags |= T3M_ZERO;
                if (s0 == 0) m.flags |= T3M_NEWCHUNK;
                if (s1 == kc) m.flags |= T3M_ENDCHUNK;
                m.wait_epi = -1;
                first = false;
                Acc a;
                a.rd.push_back({R_TMEM, m.a_col, m.a_col + 8 * m.nk});
                a.wr.push_back({R_TMEM, d_col, d_col + B.N});
                push_mma(m, a);
                s0 = s1;
            }
            k0 += kc;
        }
        if (commit) t.mmas.back().flags |= T3M_COMMIT;
        t.n_mma_instr += nk_total;
        t.tensor_cycles += (int64_t)nk_total * (B.N / 2);
    }
    void gemm_ss(int a_img, int a_rowgroup, int b_img, int N, int d_col) {
        T3Mma m{};
        m.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        m.b_off = (uint32_t)a_img | ((uint32_t)b_img << 8);
        m.d_col = (uint16_t)d_col;
        m.a_col = (uint16_t)a_rowgroup;
        m.nk = 16;
        m.b_sbo16 = 0;
        m.flags = T3M_SS | T3M_ZERO | T3M_COMMIT;
        m.wait_epi = -1;
        Acc a;
        a.rd.push_back({R_IMG0 + a_img, a_rowgroup * 8, a_rowgroup * 8 + 128});
        a.rd.push_back({R_IMG0 + b_img, 0, N});
        a.wr.push_back({R_TMEM, d_col, d_col + N});
        push_mma(m, a);
        t.n_mma_instr += 16;
        t.tensor_cycles += 16 * (N / 2);
    }
    // a record / step that writes columns a parked flush still has to read: the flush goes first (all older ones with it: in-order)
    void drain_clobbered(const Acc& a) {
        for (size_t i = pending.size(); i-- > 0;) {
            bool hit = false;
            for (const Res& w : a.wr) for (const Res& r : pending[i].second.rd) hit |= overlap(w, r);
            if (hit) { drain((int)i + 1); break; }
        }
    }
    void push_mma(const T3Mma& m, const Acc& a) {
        drain_clobbered(a);
        t.mmas.push_back(m); macc.push_back(a); mtime.push_back(clock++);
    }
    int push_epi(const T3Epi& e, const Acc& a) {
        if (e.type != T3E_FLUSH) drain_clobbered(a);
        t.epis.push_back(e); eacc.push_back(a); etime.push_back(clock++);
        return (int)t.epis.size() - 1;
    }
    int tab(const std::vector<int16_t>& v) {
        const int off = (int)t.tab16.size();
        t.tab16.insert(t.tab16.end(), v.begin(), v.end());
        return off;
    }

    // Cross-role waits from the access sets: a step waits for the latest earlier (logical time) step of the other
    // role it conflicts with.  Completion is in order inside each role, so one index per step is enough.
    void infer_waits() {
        const int nm = (int)t.mmas.size(), ne = (int)t.epis.size();
        // signal index of every MMA record = number of COMMIT flags strictly before it
        std::vector<int> sig(nm);
        int s = 0;
        for (int i = 0; i < nm; ++i) { sig[i] = s; if (t.mmas[i].flags & T3M_COMMIT) ++s; }
        t.n_mma_signals = s;
        for (int e = 0; e < ne; ++e) {
            int w = -1;
            for (int m = 0; m < nm; ++m)
                if (mtime[m] < etime[e] && conflicts(eacc[e], macc[m])) w = std::max(w, sig[m]);
            t.epis[e].wait_mma = w;
        }
        for (int m = 0; m < nm; ++m) {
            int w = -1;
            for (int e = 0; e < ne; ++e)
                if (etime[e] < mtime[m] && conflicts(macc[m], eacc[e])) w = std::max(w, e);
            t.mmas[m].wait_epi = w;
        }
        // a wait that an earlier step of the same role already implies is redundant but harmless; keep the tables simple.
        // tile boundary: the last epilogue step drains the tensor pipe, so the next tile starts from a clean state
        if (ne) t.epis[ne - 1].wait_mma = (t.n_mma_signals - 1);
    }
};

int pad8(int v) { return (v + 7) & ~7; }
int pad16(int v) { return (v + 15) & ~15; }

}  // namespace

// TMEM map of one group; false if it needs more than 512 columns.  The transport kernels need the two hidden buffers and the
// small operands only (no accumulators), which admits single nodes up to h = 232.
static bool tmem_layout(T3Group& g, bool transport) {
    const int PW = g.HP + 8;
    int c = 0;
    if (transport) {
        // first hidden layer whole (it is the K of the second), the second hidden layer whole when it fits, else (single wide
        // node) in chunks of CH columns that the third layer consumes as K chunks: h2 never exists as a whole
        const int rest = kTmemCols - PW - g.KA - g.OW - 8;
        g.CH = 0;
        if (rest < g.HP) {
            if (g.nodes.size() != 1) return false;
            g.CH = std::min(rest, 256) / 16 * 16;
            if (g.CH < 32) return false;
        }
        g.tm_p = c; c += PW;
        g.tm_q = c; c += (g.CH ? g.CH : g.HP) + 8;
        g.tm_ain = c; c += g.KA;
        g.tm_out = c; c += g.OW;
        g.tm_acc2 = g.tm_dout = g.tm_da = g.tm_acc1 = g.tm_acc3 = 0;
        return c <= kTmemCols;
    }
    g.tm_p = c; c += PW;
    g.tm_q = c; c += PW;
    g.tm_acc2 = c; c += g.N2;
    g.tm_ain = c; c += g.KA;
    g.tm_dout = c; c += g.KD;
    g.tm_out = c; c += g.OW;
    const int fixed = c;
    // preferred: separate dA, dW1 and dW3 accumulators; else aliases (the inferred waits serialise their users)
    if (fixed + 2 * g.N1 + g.OW <= kTmemCols) { g.tm_da = fixed; g.tm_acc1 = fixed + g.N1; g.tm_acc3 = fixed + 2 * g.N1; return true; }
    const int m = std::max(g.N1, g.OW);
    if (fixed + g.N1 + m <= kTmemCols) { g.tm_da = fixed; g.tm_acc1 = fixed + g.N1; g.tm_acc3 = g.tm_acc1; return true; }
    if (fixed - g.OW + 2 * m <= kTmemCols) { g.tm_da = g.tm_out; g.tm_acc1 = g.tm_out + m; g.tm_acc3 = g.tm_acc1; return true; }
    return false;
}
static void group_dims(T3Group& g, int dc) {
    g.KA = pad8(g.KX + dc + 2);   // inputs, condition, and TWO ones columns: biases enter as hi + lo tf32 parts (kT3BiasLo)
    g.KD = pad8(g.OC);
    g.OW = pad16(g.OC);
    g.N2 = pad16(g.HP + 8);
    g.N1 = pad16(g.KX + dc + 2);
    g.mtiles = (g.HP + 127) / 128;
}

void build_tc3_plan(const Plan& p, T3Plan& t, int kind) {
    t = T3Plan();
    t.kind = kind;
    t.d = p.d; t.dc = p.dc; t.alpha = p.alpha;
    const bool transport = kind != T3K_BACKWARD;
    auto fail = [&](const std::string& w) { t.ok = false; t.why = w; };

    // ---- groups: nodes of one depth, packed greedily while the group fits TMEM (and at most hp_cap hidden columns) ----
    auto build_groups = [&](int hp_cap) -> bool {
        t.groups.clear();
        for (int depth = 0; depth <= p.max_depth; ++depth) {
            T3Group g;
            for (int i = 0; i < (int)p.nodes.size(); ++i) {
                const auto& n = p.nodes[i];
                if (n.depth != depth) continue;
                const int hp = pad16(n.h);
                if (hp > (transport ? kHpMaxTransport : kHpMaxSingle)) { fail("hidden width " + std::to_string(n.h) + " exceeds the TMEM budget of the tcgen05 training kernel"); return false; }
                T3Group trial = g;
                trial.nodes.push_back(i);
                trial.hoff.push_back(g.HP); trial.xoff.push_back(g.KX); trial.ooff.push_back(g.OC);
                trial.HP += hp; trial.KX += n.k; trial.OC += n.cout;
                group_dims(trial, p.dc);
                const bool fits = trial.nodes.size() == 1 ||
                                  (trial.HP <= hp_cap && trial.KA <= (transport ? kKaMaxTransport : kKaMax) && trial.OW <= (transport ? kOwMaxTransport : kOwMax) &&
                                   (int)trial.nodes.size() <= 16 && tmem_layout(trial, transport));
                if (!fits) {
                    t.groups.push_back(g);
                    trial = T3Group();
                    trial.nodes.push_back(i);
                    trial.hoff.push_back(0); trial.xoff.push_back(0); trial.ooff.push_back(0);
                    trial.HP = hp; trial.KX = n.k; trial.OC = n.cout;
                    group_dims(trial, p.dc);
                }
                g = trial;
            }
            if (!g.nodes.empty()) t.groups.push_back(g);
        }
        for (T3Group& g : t.groups) {
            group_dims(g, p.dc);
            if (g.KA > (transport ? kKaMaxTransport : kKaMax) || g.OW > (transport ? kOwMaxTransport : kOwMax)) { fail("a node's input/output width exceeds the tcgen05 training kernel's envelope"); return false; }
            if (!tmem_layout(g, transport)) { fail("a tree level does not fit the 512 TMEM columns"); return false; }
        }
        return true;
    };

    // ---- shared memory map ----
    t.xp = (p.d + p.dc) | 1;
    int HPmax = 0, N1max = 0, N2max = 0, OCmax = 0, OWmax = 0, KAmax = 0, KDmax = 0, mtmax = 1, tab_bytes = 0, epi_bytes = 0;
    auto layout = [&](int nimg, int nslots, int slot) {
        int o = 0;
        t.sm_bars = o; o += 1024;
        for (int i = 0; i < kT3Imgs; ++i) { t.sm_img[i] = 0; t.img_rows[i] = 0; }
        // image 0 (h1) is also the B operand of dW2 (N2 rows incl. the ones block); images 1, 2 are A operands only.
        // An M tile reads 128 rows from its first row group: rows past an image land in the regions that follow it
        // (finite or not, they only reach accumulator rows that are never flushed) - the tail check below keeps
        // those reads inside the CTA's allocation.
        // Rows are allocated for what the epilogue WRITES (hidden features + the ones block, inputs, outputs), not for the MMA's
        // N (a multiple of 16): as a B operand an image is read for N rows from its first row, and the rows past the allocation
        // land in the regions that follow it - garbage that only reaches accumulator columns no parameter maps to.  The shared
        // memory saved this way pays for 32 KB weight slabs: the single issuer thread spends ~450 cycles of its own per record
        // (barrier poll, descriptor arithmetic, two commits) against 65 cycles of tensor time per MMA, so a record must carry
        // 8 MMAs, not 4, to keep the pipe busy.
        for (int i = 0; i < nimg; ++i) {
            t.sm_img[i] = o; t.img_rows[i] = i == 0 ? pad8(HPmax + 8) : pad8(HPmax);
            o += t.img_rows[i] * 512;
        }
        if (!transport) {     // the transport programs write no images at all (T3E_IN carries T3I_NOIMG)
            t.sm_img[3] = o; t.img_rows[3] = pad8(KAmax); o += t.img_rows[3] * 512;
            t.sm_img[4] = o; t.img_rows[4] = pad8(KDmax); o += t.img_rows[4] * 512;
        }
        t.sm_ring = o; o += nslots * slot;
        t.sm_tab16 = o; o += tab_bytes;
        t.sm_epis = o; o += epi_bytes;
        t.sm_xs = o; o += 128 * t.xp * 4;
        t.sm_gs = o; o += transport ? 0 : 128 * t.xp * 4;     // gradient state: backward only
        t.sm_os = o; o += 128 * t.op * 4;
        t.sm_stage = o; o += 0;
        t.sm_red = o; o += 4 * 32 * 4;
        // tail: the furthest byte an over-reading M tile (A operand, 128 rows per tile) or B operand (N rows) can touch
        for (int i = 0; i < nimg; ++i) {
            const int last = t.sm_img[i] + 3 * t.img_rows[i] * 128 + (mtmax * 128) * 128;
            o = std::max(o, last);
        }
        if (!transport) {
            o = std::max(o, t.sm_img[0] + 3 * t.img_rows[0] * 128 + N2max * 128);
            o = std::max(o, t.sm_img[3] + 3 * t.img_rows[3] * 128 + N1max * 128);
            o = std::max(o, t.sm_img[4] + 3 * t.img_rows[4] * 128 + OWmax * 128);
        }
        t.n_imgs_hidden = nimg; t.n_slots = nslots; t.slot_bytes = slot;
        t.smem_bytes = (o + 15) & ~15;
        return t.smem_bytes <= kSmemMax - 1024;   // 1 KB left for the kernel's static shared memory
    };
    bool placed = false;
    // backward: a multi-node group is one M tile of the weight-gradient GEMMs (128 hidden columns); the transport programs have no
    // such limit, only the TMEM budget
    const std::vector<int> caps = transport ? std::vector<int>{224, 192, 160, 128, 112, 96, 80, 64, 48, 32} : std::vector<int>{128, 112, 96, 80, 64, 48, 32};
    for (int hp_cap : caps) {
        if (!build_groups(hp_cap)) return;   // `why` already set: no cap helps
        tab_bytes = 0; epi_bytes = 0;
        for (const T3Group& g : t.groups) {
            tab_bytes += 2 * (g.KA + g.OC + g.KX + p.dc + 1 + 6 * (int)g.nodes.size());
            epi_bytes += (int)sizeof(T3Epi) * (16 + 6 * g.mtiles + (g.CH ? 4 * ((g.HP + g.CH - 1) / g.CH) : 0));   // steps of one group: 4 + 9 + 9, flushes of extra M tiles, hidden chunks
        }
        tab_bytes = (tab_bytes + 15) & ~15;
        HPmax = N1max = N2max = OCmax = OWmax = KAmax = KDmax = 0; mtmax = 1;
        for (const T3Group& g : t.groups) {
            KAmax = std::max(KAmax, g.KA); KDmax = std::max(KDmax, g.KD);
            HPmax = std::max(HPmax, g.HP); N1max = std::max(N1max, g.N1); N2max = std::max(N2max, g.N2);
            OCmax = std::max(OCmax, g.OC); OWmax = std::max(OWmax, g.OW); mtmax = std::max(mtmax, g.mtiles);
        }
        t.op = OCmax | 1;
        if (transport) {
            if (layout(0, 4, 36864) || layout(0, 3, 36864) || layout(0, 4, 32768) || layout(0, 3, 32768) || layout(0, 4, 16384) || layout(0, 3, 16384) || layout(0, 2, 16384) || layout(0, 2, 8192)) { placed = true; break; }
            continue;
        }
        if (layout(2, 2, 36864) || layout(2, 2, 32768) || layout(3, 3, 16384) || layout(2, 3, 16384) || layout(2, 2, 16384) || layout(2, 3, 8192) || layout(2, 2, 8192)) { placed = true; break; }
    }
    if (!placed) return fail("the block's widest level does not fit shared memory");

    // ---- partial-gradient layout + unpack map ----
    t.unpack_src.assign((size_t)p.n_params, -1);
    t.unpack_q4.assign((size_t)p.n_params, 0);
    if (!transport) {
        int64_t o = 0;
        for (T3Group& g : t.groups) {
            const int rows = g.mtiles * 128;
            for (int net = 0; net < 2; ++net) {
                g.part[net][0] = (int)o; o += (int64_t)rows * g.N2;
                g.part[net][1] = (int)o; o += (int64_t)rows * g.N1;
                g.part[net][2] = (int)o; o += (int64_t)rows * g.OW;
                g.part[net][3] = (int)o; o += 4 * 32;   // layer-3 bias gradients: one slot per lane quadrant (deterministic, no atomics)
                // accumulator blocks are stored lane-fastest ([M tile][column][128 lanes]): the flush (thread = lane) is coalesced
                auto at = [&](int blk, int N, int row, int col) { return g.part[net][blk] + (row / 128) * (N * 128) + col * 128 + (row % 128); };
                for (size_t q = 0; q < g.nodes.size(); ++q) {
                    const int ni = g.nodes[q];
                    const auto& n = p.nodes[ni];
                    const int ho = g.hoff[q], xo = g.xoff[q], oo = g.ooff[q];
                    for (int j = 0; j < n.h; ++j) {
                        for (int m = 0; m < n.k; ++m)
                            t.unpack_src[(size_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + m)] = at(1, g.N1, ho + j, xo + m);
                        for (int q2 = 0; q2 < p.dc; ++q2)
                            t.unpack_src[(size_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + n.k + q2)] = at(1, g.N1, ho + j, g.KX + q2);
                        t.unpack_src[(size_t)(poffc(p, ni, net, 0, 1) + j)] = at(1, g.N1, ho + j, g.KX + p.dc);
                        for (int m = 0; m < n.h; ++m)
                            t.unpack_src[(size_t)(poffc(p, ni, net, 1, 0) + (int64_t)j * n.h + m)] = at(0, g.N2, ho + j, ho + m);
                        t.unpack_src[(size_t)(poffc(p, ni, net, 1, 1) + j)] = at(0, g.N2, ho + j, g.HP);
                    }
                    for (int c = 0; c < n.cout; ++c) {
                        for (int m = 0; m < n.h; ++m)
                            t.unpack_src[(size_t)(poffc(p, ni, net, 2, 0) + (int64_t)c * n.h + m)] = at(2, g.OW, ho + m, oo + c);
                        t.unpack_src[(size_t)(poffc(p, ni, net, 2, 1) + c)] = g.part[net][3] + oo + c;
                        t.unpack_q4[(size_t)(poffc(p, ni, net, 2, 1) + c)] = 1;
                    }
                }
            }
        }
        t.n_partial = o;
        for (int32_t v : t.unpack_src) if (v < 0) return fail("internal: parameter without a partial-gradient slot");
    }

    // ---- programs ----
    Builder b(p, t);
    const bool defer_flush = true;
    const int IMG_IN = 3, IMG_DOUT = 4;
    const int IMG_H1 = 0, IMG_H2 = 1, IMG_DH2 = t.n_imgs_hidden == 3 ? 2 : 1, IMG_DH1 = 1;
    // group order: backward and inverse visit the root level first (hint.py:85-88), forward the deepest level first (hint.py:70-73)
    std::vector<size_t> gorder(t.groups.size());
    for (size_t i = 0; i < gorder.size(); ++i) gorder[i] = kind == T3K_FORWARD ? gorder.size() - 1 - i : i;
    for (size_t gi : gorder) {
        T3Group& g = t.groups[gi];
        const int nn = (int)g.nodes.size();
        const int P = g.tm_p, Q = g.tm_q;
        // tables of this group
        std::vector<int16_t> in_src(g.KA, -1), out_x(g.OC, 0), da_dst(g.KX + p.dc, 0), ntab;
        ntab.push_back(nn);
        for (int q = 0; q < nn; ++q) {
            const auto& n = p.nodes[g.nodes[q]];
            for (int m = 0; m < n.k; ++m) { in_src[g.xoff[q] + m] = (n.lo + m); da_dst[g.xoff[q] + m] = (n.lo + m); }
            for (int c = 0; c < n.cout; ++c) out_x[g.ooff[q] + c] = (n.lo + n.k + c);
            ntab.push_back(g.hoff[q]); ntab.push_back(n.h);
            ntab.push_back(g.xoff[q]); ntab.push_back(n.k);
            ntab.push_back(g.ooff[q]); ntab.push_back(n.cout);
        }
        for (int q2 = 0; q2 < p.dc; ++q2) { in_src[g.KX + q2] = (p.d + q2); da_dst[g.KX + q2] = (p.d + q2); }
        in_src[g.KX + p.dc] = -2;
        in_src[g.KX + p.dc + 1] = -2;
        const int tab_in = b.tab(in_src), tab_out = b.tab(out_x), tab_da = b.tab(da_dst);
        g.tab_nodes = b.tab(ntab);

        // weight operands of the group, per net
        auto w1g = [&](int net) {
            BOp B; B.init(g.HP, g.KA);
            for (int q = 0; q < nn; ++q) {
                const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
                for (int j = 0; j < n.h; ++j) {
                    for (int m = 0; m < n.k; ++m) B.at(g.hoff[q] + j, g.xoff[q] + m) = (int32_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + m);
                    for (int q2 = 0; q2 < p.dc; ++q2) B.at(g.hoff[q] + j, g.KX + q2) = (int32_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + n.k + q2);
                    B.at(g.hoff[q] + j, g.KX + p.dc) = (int32_t)(poffc(p, ni, net, 0, 1) + j);
                    B.at(g.hoff[q] + j, g.KX + p.dc + 1) = (int32_t)(poffc(p, ni, net, 0, 1) + j) | kT3BiasLo;
                }
            }
            return B;
        };
        auto w2n = [&](int net, int q) {   // [pad16(h)][pad8(h) + 8]: last K step = bias against the ones block
            const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
            BOp B; B.init(pad16(n.h), pad8(n.h) + 8);
            for (int j = 0; j < n.h; ++j) {
                for (int m = 0; m < n.h; ++m) B.at(j, m) = (int32_t)(poffc(p, ni, net, 1, 0) + (int64_t)j * n.h + m);
                B.at(j, pad8(n.h)) = (int32_t)(poffc(p, ni, net, 1, 1) + j);
                B.at(j, pad8(n.h) + 1) = (int32_t)(poffc(p, ni, net, 1, 1) + j) | kT3BiasLo;
            }
            return B;
        };
        auto w3g = [&](int net) {           // [OW][HP + 8]
            BOp B; B.init(g.OW, g.HP + 8);
            for (int q = 0; q < nn; ++q) {
                const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
                for (int c = 0; c < n.cout; ++c) {
                    for (int m = 0; m < n.h; ++m) B.at(g.ooff[q] + c, g.hoff[q] + m) = (int32_t)(poffc(p, ni, net, 2, 0) + (int64_t)c * n.h + m);
                    B.at(g.ooff[q] + c, g.HP) = (int32_t)(poffc(p, ni, net, 2, 1) + c);
                    B.at(g.ooff[q] + c, g.HP + 1) = (int32_t)(poffc(p, ni, net, 2, 1) + c) | kT3BiasLo;
                }
            }
            return B;
        };
        auto w3tg = [&](int net) {          // dH2 = dOut * W3 : [HP][KD]
            BOp B; B.init(g.HP, g.KD);
            for (int q = 0; q < nn; ++q) {
                const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
                for (int c = 0; c < n.cout; ++c)
                    for (int m = 0; m < n.h; ++m) B.at(g.hoff[q] + m, g.ooff[q] + c) = (int32_t)(poffc(p, ni, net, 2, 0) + (int64_t)c * n.h + m);
            }
            return B;
        };
        auto w2tn = [&](int net, int q) {   // dH1 = dH2 * W2 : [pad16(h)][pad8(h)]
            const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
            BOp B; B.init(pad16(n.h), pad8(n.h));
            for (int j = 0; j < n.h; ++j)
                for (int m = 0; m < n.h; ++m) B.at(m, j) = (int32_t)(poffc(p, ni, net, 1, 0) + (int64_t)j * n.h + m);
            return B;
        };
        auto w1tg = [&](int net) {          // dA = dH1 * W1 : [N1][HP]
            BOp B; B.init(g.N1, g.HP);
            for (int q = 0; q < nn; ++q) {
                const int ni = g.nodes[q]; const auto& n = p.nodes[ni];
                for (int j = 0; j < n.h; ++j) {
                    for (int m = 0; m < n.k; ++m) B.at(g.xoff[q] + m, g.hoff[q] + j) = (int32_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + m);
                    for (int q2 = 0; q2 < p.dc; ++q2) B.at(g.KX + q2, g.hoff[q] + j) = (int32_t)(poffc(p, ni, net, 0, 0) + (int64_t)j * n.cin + n.k + q2);
                }
            }
            return B;
        };

        // ---- emitters ----
        auto e_in = [&]() {
            T3Epi e{}; e.type = T3E_IN; e.a = tab_in; e.b = g.KA; e.c = g.tm_ain; e.flags = transport ? T3I_NOIMG : 0;
            Acc a; a.wr.push_back({R_TMEM, g.tm_ain, g.tm_ain + g.KA}); if (!transport) a.wr.push_back({R_IMG0 + IMG_IN, 0, g.KA});
            b.push_epi(e, a);
        };
        auto e_hid = [&](int col0, int img, bool img_ones, int width = -1, bool ones = true) {   // img < 0: no image
            const int w = width < 0 ? g.HP : width;
            T3Epi e{}; e.type = T3E_HID; e.flags = ((ones ? T3H_ONES : 0) | (img >= 0 ? T3H_IMG : 0) | (img >= 0 && img_ones ? T3H_IMG_ONES : 0));
            e.a = col0; e.b = w; e.c = (img < 0 ? 0 : img);
            Acc a; a.rd.push_back({R_TMEM, col0, col0 + w}); a.wr.push_back({R_TMEM, col0, col0 + w + (ones ? 8 : 0)});
            if (img >= 0) a.wr.push_back({R_IMG0 + img, 0, w + (img_ones ? 8 : 0)});
            b.push_epi(e, a);
        };
        // rows [n0, n1) / columns [k0, k1) of an operand
        auto rows_of = [&](const BOp& B, int n0, int n1) {
            BOp R; R.init(n1 - n0, B.K);
            for (int n = n0; n < n1; ++n) for (int k = 0; k < B.K; ++k) R.at(n - n0, k) = B.src[(size_t)n * B.K + k];
            return R;
        };
        auto cols_of = [&](const BOp& B, int k0, int k1) {
            BOp R; R.init(B.N, k1 - k0);
            for (int n = 0; n < B.N; ++n) for (int k = k0; k < k1; ++k) R.at(n, k - k0) = B.src[(size_t)n * B.K + k];
            return R;
        };
        auto e_flush = [&](int net, int kind, int mt) {
            const int col0 = kind == T3F_W2 ? g.tm_acc2 : kind == T3F_W1 ? g.tm_acc1 : g.tm_acc3;
            const int N = kind == T3F_W2 ? g.N2 : kind == T3F_W1 ? g.N1 : g.OW;
            const int blk = kind == T3F_W2 ? 0 : kind == T3F_W1 ? 1 : 2;
            T3Epi e{}; e.type = T3E_FLUSH; e.a = col0; e.b = N; e.off = g.part[net][blk] + mt * 128 * N; e.d = 128;
            e.e = std::min(128, g.HP - mt * 128); e.f = g.tab_nodes; e.g = kind; e.h = mt;
            e.c = (kind == T3F_W2 ? g.HP : kind == T3F_W1 ? g.KX : -1);   // first "extra" column (bias / condition block)
            Acc a; a.rd.push_back({R_TMEM, col0, col0 + N});
            if (defer_flush && kind != T3F_W3) b.defer(e, a); else b.push_epi(e, a);   // the dW3 flush already runs behind the dH1 GEMM
        };
        auto fwd_chain = [&](int net, bool images, bool layer3) {
            if (g.CH) {
                // wide single node (transport programs only): layer 1 in N halves (MMA N <= 256); layer 2 in column chunks of CH
                // whose relu'd result is the K chunk of a layer-3 partial product - h2 never exists as a whole
                const auto& n = p.nodes[g.nodes[0]];
                const BOp W1 = w1g(net), W2 = w2n(net, 0), W3 = w3g(net);
                for (int n0 = 0; n0 < g.HP; n0 += 256) {
                    const int n1 = std::min(g.HP, n0 + 256);
                    b.gemm_ts(rows_of(W1, n0, n1), {{g.tm_ain, g.KA / 8}}, P + n0, n1 == g.HP);
                }
                e_hid(P, -1, false);
                const int K2 = pad8(n.h);
                for (int j0 = 0; j0 < g.HP; j0 += g.CH) {
                    const int j1 = std::min(g.HP, j0 + g.CH), w = j1 - j0;
                    const bool last = j1 == g.HP;
                    b.gemm_ts(rows_of(W2, j0, j1), {{P, K2 / 8}, {P + g.HP, 1}}, Q, true);
                    e_hid(Q, -1, false, w, last);
                    BOp W3c = cols_of(W3, j0, last ? g.HP + 8 : j1);      // the last chunk carries the bias K step
                    std::vector<ASeg> as{{Q, w / 8}};
                    if (last) as.push_back({Q + w, 1});
                    b.gemm_ts(W3c, as, g.tm_out, last, j0 == 0);
                }
                return;
            }
            b.gemm_ts(w1g(net), {{g.tm_ain, g.KA / 8}}, P, true);
            e_hid(P, images ? IMG_H1 : -1, true);
            for (int q = 0; q < nn; ++q) {
                const auto& n = p.nodes[g.nodes[q]];
                b.gemm_ts(w2n(net, q), {{P + g.hoff[q], pad8(n.h) / 8}, {P + g.HP, 1}}, Q + g.hoff[q], q == nn - 1);
            }
            b.drain(-1);   // parked flushes of the previous backward chain run behind the hidden-layer GEMM just issued
            e_hid(Q, images ? IMG_H2 : -1, false);
            if (layer3) b.gemm_ts(w3g(net), {{Q, g.HP / 8 + 1}}, g.tm_out, true);
        };
        auto bwd_chain = [&](int net) {
            // dH2 = dOut * W3, masked by h2 > 0 (h2 still in Q)
            b.gemm_ts(w3tg(net), {{g.tm_dout, g.KD / 8}}, P, true);
            // dW3^T [h2 feature][out] = H2^T dOut ; the dH2 epilogue overlaps the last M tile's GEMM when its image does
            // not alias the h2 image
            for (int mt = 0; mt < g.mtiles; ++mt) {
                b.gemm_ss(IMG_H2, mt * 16, IMG_DOUT, g.OW, g.tm_acc3);
                if (mt == g.mtiles - 1) {
                    T3Epi d{}; d.type = T3E_DHID; d.flags = T3D_MASK_TMEM; d.a = P; d.b = g.HP; d.c = IMG_DH2; d.e = Q;
                    Acc da; da.rd.push_back({R_TMEM, P, P + g.HP}); da.rd.push_back({R_TMEM, Q, Q + g.HP});
                    da.wr.push_back({R_TMEM, P, P + g.HP}); da.wr.push_back({R_IMG0 + IMG_DH2, 0, g.HP});
                    b.push_epi(d, da);
                }
                e_flush(net, T3F_W3, mt);
            }
            // dH1 = dH2 * W2 per node, masked by h1 > 0 (h1 image)
            for (int q = 0; q < nn; ++q) {
                const auto& n = p.nodes[g.nodes[q]];
                b.gemm_ts(w2tn(net, q), {{P + g.hoff[q], pad8(n.h) / 8}}, Q + g.hoff[q], q == nn - 1);
            }
            // dW2 [h2 feature][h1 feature | 1] = dH2^T [H1 | 1]
            for (int mt = 0; mt < g.mtiles; ++mt) {
                b.gemm_ss(IMG_DH2, mt * 16, IMG_H1, g.N2, g.tm_acc2);
                if (mt == g.mtiles - 1) {
                    T3Epi d{}; d.type = T3E_DHID; d.flags = 0; d.a = Q; d.b = g.HP; d.c = IMG_DH1; d.e = IMG_H1;
                    Acc da; da.rd.push_back({R_TMEM, Q, Q + g.HP}); da.rd.push_back({R_IMG0 + IMG_H1, 0, g.HP});
                    da.wr.push_back({R_TMEM, Q, Q + g.HP}); da.wr.push_back({R_IMG0 + IMG_DH1, 0, g.HP});
                    b.push_epi(d, da);
                }
                e_flush(net, T3F_W2, mt);
            }
            // dA = dH1 * W1
            b.gemm_ts(w1tg(net), {{Q, g.HP / 8}}, g.tm_da, true);
            {
                T3Epi e{}; e.type = T3E_DA; e.a = g.tm_da; e.b = (g.KX + p.dc); e.c = tab_da;
                Acc a; a.rd.push_back({R_TMEM, g.tm_da, g.tm_da + g.KX + p.dc});
                b.push_epi(e, a);
            }
            // dW1 [h1 feature][input | 1] = dH1^T [A | 1]
            for (int mt = 0; mt < g.mtiles; ++mt) {
                b.gemm_ss(IMG_DH1, mt * 16, IMG_IN, g.N1, g.tm_acc1);
                e_flush(net, T3F_W1, mt);
            }
        };

        // ---- phase 1: S forward, keep s ----
        e_in();
        fwd_chain(0, false, true);
        {
            T3Epi e{}; e.type = T3E_OUTS; e.a = g.tm_out; e.b = g.OC;
            Acc a; a.rd.push_back({R_TMEM, g.tm_out, g.tm_out + g.OC});
            b.push_epi(e, a);
        }
        if (transport) {
            // ---- transport kernels: T forward, then the coupling itself (x_l <- e(s) x_l + t  or  (x_l - t) / e(s)) ----
            fwd_chain(1, false, true);
            T3Epi e{}; e.type = T3E_CPLF; e.a = g.tm_out; e.b = g.OC; e.c = tab_out; e.flags = kind == T3K_INVERSE ? 1 : 0;
            Acc a; a.rd.push_back({R_TMEM, g.tm_out, g.tm_out + g.OC});
            b.push_epi(e, a);
            continue;
        }
        // ---- phase 2: T forward, coupling, T backward ----
        fwd_chain(1, true, true);
        {
            T3Epi e{}; e.type = T3E_CPL; e.a = g.tm_out; e.b = g.OC; e.c = tab_out; e.d = g.tm_dout; e.e = g.KD;
            e.off = g.part[1][3];
            Acc a; a.rd.push_back({R_TMEM, g.tm_out, g.tm_out + g.OC});
            a.wr.push_back({R_TMEM, g.tm_dout, g.tm_dout + g.KD}); a.wr.push_back({R_IMG0 + IMG_DOUT, 0, g.OC});
            b.push_epi(e, a);
        }
        bwd_chain(1);
        // ---- phase 3: S forward again (activations), S backward ----
        fwd_chain(0, true, false);
        {
            T3Epi e{}; e.type = T3E_DS; e.a = g.tm_dout; e.b = g.OC; e.e = g.KD; e.off = g.part[0][3];
            Acc a; a.wr.push_back({R_TMEM, g.tm_dout, g.tm_dout + g.KD}); a.wr.push_back({R_IMG0 + IMG_DOUT, 0, g.OC});
            b.push_epi(e, a);
        }
        bwd_chain(0);
    }
    b.drain(-1);
    b.infer_waits();
    t.n_packed = (int64_t)t.pack_src.size();
    if ((int)t.tab16.size() * 2 > tab_bytes || (int)(t.epis.size() * sizeof(T3Epi)) > epi_bytes) return fail("internal: epilogue tables exceed their shared-memory reservation");
    for (const T3Chunk& c : t.chunks)
        if ((int)c.bytes > t.slot_bytes) return fail("internal: weight slab larger than a ring slot");
    if (t.epis.size() > 32000 || t.mmas.size() > 32000) return fail("program too long");
    t.ok = true;
}

}  // namespace hint
