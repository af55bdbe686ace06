// Planner of the warp-MMA kernels (host only, no CUDA): operand packing in B-fragment order, the partial-gradient
// layout, the shared-memory column allocation per stage and the per-warp task lists of every phase.
// Tree rules come from plan.cpp (hint.py:25-54); per-node math from hint.py:62-101 (see mma_kernels.cuh).
#include "plan_mma.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace hint {

namespace {

int64_t poff(const Plan& p, int node, int net, int layer, int kind) {
    return p.param_offsets[(size_t)node * 12 + net * 6 + layer * 2 + kind];
}

// float index of B[k][n] inside a fragment-ordered operand of NT n-tiles: k-step major, then n-tile, then lane
// (g = n%8, t = (k%8)/2), then the two k-slots of the lane (features 2t, 2t+1 of the group of 8)
inline int64_t frag(int NT, int k, int n) {
    const int ks = k >> 3, kk = k & 7, j = n >> 3, g = n & 7, t = kk >> 1;
    return ((int64_t)(ks * NT + j) * 32 + g * 4 + t) * 2 + (kk & 1);
}

struct NodeOps {       // packed offsets of one node
    int w1, b1;        // L1, s and t fused on N: K = cin, N = 2*h8
    int w2[2], b2[2];  // K = h8, N = h8
    int w3[2], b3[2];  // K = h8, N = c8
    int g3[2];         // dH2 = dOut * W3      : K = c8, N = h8
    int g2[2];         // dH1 = dH2 * W2       : K = h8, N = h8
    int g1;            // dx_upper = dH1 * W1x : K = 2*h8 (s then t hidden units), N = round8(k)
    int dw[2][3];      // partial-gradient offsets: dW [N][ld], ld = round8(K+1), bias gradient in column K
    int ld[2][3];
};

struct Packer {
    MmaPlan& m;
    int64_t off = 0;
    int alloc(int64_t n) {
        const int64_t o = off;
        off += (n + 3) & ~int64_t(3);
        m.pack_src.resize((size_t)off, -1);
        return (int)o;
    }
    void set(int64_t i, int64_t src) { m.pack_src[(size_t)i] = (int32_t)src; }
    void set_exact(int64_t i, int64_t src) { m.pack_src[(size_t)i] = (int32_t)(-src - 2); }   // biases: not rounded
};

void layer_dims(const hint_node_info_t& n, int layer, int& K, int& N) {
    if (layer == 0) { K = n.cin; N = n.h; }
    else if (layer == 1) { K = n.h; N = n.h; }
    else { K = n.h; N = n.cout; }
}

void pack_nodes(const Plan& p, MmaPlan& m, std::vector<NodeOps>& ops) {
    Packer pk{m};
    ops.assign(p.nodes.size(), NodeOps{});
    int64_t poffp = 0;
    m.unpack_src.assign((size_t)p.n_params, -1);
    for (size_t i = 0; i < p.nodes.size(); ++i) {
        const auto& n = p.nodes[i];
        NodeOps& o = ops[i];
        const int h8 = round8(n.h), c8 = round8(n.cout), k8 = round8(std::max(n.k, 1)), cin8 = round8(n.cin);
        const int ni = (int)i;
        // L1 fused
        o.w1 = pk.alloc((int64_t)cin8 * 2 * h8);
        o.b1 = pk.alloc(2 * h8);
        for (int net = 0; net < 2; ++net) {
            const int64_t w = poff(p, ni, net, 0, 0), b = poff(p, ni, net, 0, 1);
            for (int u = 0; u < n.h; ++u) {
                for (int k = 0; k < n.cin; ++k) pk.set(o.w1 + frag(2 * h8 / 8, k, net * h8 + u), w + (int64_t)u * n.cin + k);
                pk.set_exact(o.b1 + net * h8 + u, b + u);
            }
        }
        for (int net = 0; net < 2; ++net) {
            const int64_t w1 = poff(p, ni, net, 0, 0);
            const int64_t w2 = poff(p, ni, net, 1, 0), b2 = poff(p, ni, net, 1, 1);
            const int64_t w3 = poff(p, ni, net, 2, 0), b3 = poff(p, ni, net, 2, 1);
            o.w2[net] = pk.alloc((int64_t)h8 * h8);
            o.b2[net] = pk.alloc(h8);
            o.w3[net] = pk.alloc((int64_t)h8 * c8);
            o.b3[net] = pk.alloc(c8);
            o.g3[net] = pk.alloc((int64_t)c8 * h8);
            o.g2[net] = pk.alloc((int64_t)h8 * h8);
            for (int u = 0; u < n.h; ++u) {          // u = output unit of layer 2
                for (int k = 0; k < n.h; ++k) {
                    pk.set(o.w2[net] + frag(h8 / 8, k, u), w2 + (int64_t)u * n.h + k);   // B[k][u] = W2[u][k]
                    pk.set(o.g2[net] + frag(h8 / 8, u, k), w2 + (int64_t)u * n.h + k);   // B[u][k] = W2[u][k]
                }
                pk.set_exact(o.b2[net] + u, b2 + u);
            }
            for (int r = 0; r < n.cout; ++r) {
                for (int k = 0; k < n.h; ++k) {
                    pk.set(o.w3[net] + frag(c8 / 8, k, r), w3 + (int64_t)r * n.h + k);   // B[k][r] = W3[r][k]
                    pk.set(o.g3[net] + frag(h8 / 8, r, k), w3 + (int64_t)r * n.h + k);   // B[r][k] = W3[r][k]
                }
                pk.set_exact(o.b3[net] + r, b3 + r);
            }
            (void)w1;
        }
        o.g1 = pk.alloc((int64_t)2 * h8 * k8);
        for (int net = 0; net < 2; ++net) {
            const int64_t w1 = poff(p, ni, net, 0, 0);
            for (int u = 0; u < n.h; ++u)
                for (int j = 0; j < n.k; ++j) pk.set(o.g1 + frag(k8 / 8, net * h8 + u, j), w1 + (int64_t)u * n.cin + j);
        }
        // partial-gradient layout
        for (int net = 0; net < 2; ++net)
            for (int layer = 0; layer < 3; ++layer) {
                int K, N;
                layer_dims(n, layer, K, N);
                const int ld = round8(K + 1);
                o.dw[net][layer] = (int)poffp;
                o.ld[net][layer] = ld;
                const int64_t w = poff(p, ni, net, layer, 0), b = poff(p, ni, net, layer, 1);
                for (int r = 0; r < N; ++r) {
                    for (int k = 0; k < K; ++k) m.unpack_src[(size_t)(w + (int64_t)r * K + k)] = (int32_t)(poffp + (int64_t)r * ld + k);
                    m.unpack_src[(size_t)(b + r)] = (int32_t)(poffp + (int64_t)r * ld + K);
                }
                poffp += (int64_t)N * ld;
                poffp = (poffp + 3) & ~int64_t(3);
            }
    }
    m.n_packed = pk.off;
    m.n_partial = poffp;
}

// ---- task generation ---------------------------------------------------------------------------------------------
struct FJob {   // forward-type GEMM job before chunking
    int w_off, NT, b_off, in0, k0, in1, k1, ksteps, out0, nvalid, flags;
    bool raw_input = false;   // reads the exact x / condition columns (layer 1): never FAST
};
struct DJob {
    int n_col, N, in0, k0, in1, k1, out_off, ld;
};
struct Cand {
    std::vector<MTask> mt;
    std::vector<DTask> dt;
    std::vector<double> mcost, dcost;
};

void chunk_sizes(int total, int cmax, std::vector<int>& out) {
    out.clear();
    const int nch = (total + cmax - 1) / cmax;
    for (int i = 0; i < nch; ++i) out.push_back(total / nch + (i < total % nch ? 1 : 0));
}

void gen_ftasks(const std::vector<FJob>& jobs, const MSchedule& s, int cmax, Cand& c) {
    const int TMS = s.TM + 4, mt = s.TM / 16;
    std::vector<int> cs;
    for (const FJob& j : jobs) {
        chunk_sizes(j.NT, cmax, cs);
        int j0 = 0;
        for (int nt : cs) {
            MTask t{};
            t.w_off = j.w_off + j0 * 64;
            t.ks_stride = (unsigned short)(j.NT * 64);
            t.b_off = j.b_off < 0 ? -1 : j.b_off + j0 * 8;
            t.in_off = (unsigned short)(j.in0 * TMS); t.k0 = (unsigned short)j.k0;
            t.in1_off = (unsigned short)(j.in1 * TMS); t.k1 = (unsigned short)j.k1;
            t.ksteps = (unsigned short)j.ksteps;
            t.out_off = (unsigned short)((j.out0 + j0 * 8) * TMS);
            t.zero_off = (unsigned short)(s.col_zero * TMS);
            const int nvalid = std::max(0, std::min(nt * 8, j.nvalid - j0 * 8));
            t.nvalid = (unsigned short)nvalid;
            t.mt = (unsigned char)mt; t.nt = (unsigned char)nt;
            t.flags = (unsigned char)(j.flags | ((!j.raw_input && j.k1 == 0 && j.k0 == 8 * j.ksteps) ? MT_FAST : 0));
            t.type = OP_GEMM;
            if (nvalid > 0) {
                c.mt.push_back(t);
                c.mcost.push_back(j.ksteps * (mt * nt * 8.0 + 4.0 * mt + 12.0 * nt + 16.0) + 12.0 * mt * nt + 60.0);
            }
            j0 += nt;
        }
    }
}

void gen_dtasks(const std::vector<DJob>& jobs, const MSchedule& s, int mt, int cmax, Cand& c) {
    const int TM = s.TM, TMS = s.TM + 4;
    std::vector<int> cs;
    for (const DJob& j : jobs) {
        const int MTt = (j.N + 15) / 16, NTt = (j.k0 + j.k1 + 1 + 7) / 8;
        chunk_sizes(NTt, cmax, cs);
        for (int i0 = 0; i0 < MTt; i0 += mt) {
            const int mm = std::min(mt, MTt - i0);
            int j0 = 0;
            for (int nt : cs) {
                DTask t{};
                t.a_off = (unsigned short)(j.n_col * TMS); t.N = (unsigned short)j.N; t.n0 = (unsigned short)(16 * i0);
                t.in_off = (unsigned short)(j.in0 * TMS); t.k0 = (unsigned short)j.k0;
                t.in1_off = (unsigned short)(j.in1 * TMS); t.k1 = (unsigned short)j.k1;
                t.kf0 = (unsigned short)(8 * j0);
                t.out_off = j.out_off; t.ld = (unsigned short)j.ld;
                t.mt = (unsigned char)mm; t.nt = (unsigned char)nt;
                t.nstore = (unsigned short)std::min(j.ld, j.k0 + j.k1 + 1);
                t.one_off = (unsigned short)(s.col_one * TMS); t.zero_off = (unsigned short)(s.col_zero * TMS);
                t.type = OP_DW;
                c.dt.push_back(t);
                c.dcost.push_back((TM / 8) * (mm * nt * 8.0 + 4.0 * mm + 2.0 * nt + 16.0) + 30.0 * mm * nt + 150.0);
                j0 += nt;
            }
        }
    }
}

// LPT assignment of the candidate's tasks to the warps; returns the makespan and the per-task warp
double assign(const Cand& c, std::vector<int>& mw, std::vector<int>& dw) {
    struct It { double cost; int kind, idx; };
    std::vector<It> items;
    for (size_t i = 0; i < c.mt.size(); ++i) items.push_back({c.mcost[i], 0, (int)i});
    for (size_t i = 0; i < c.dt.size(); ++i) items.push_back({c.dcost[i], 1, (int)i});
    std::stable_sort(items.begin(), items.end(), [](const It& a, const It& b) { return a.cost > b.cost; });
    double load[kMmaWarps] = {0};
    mw.assign(c.mt.size(), 0);
    dw.assign(c.dt.size(), 0);
    for (const It& it : items) {
        int best = 0;
        for (int w = 1; w < kMmaWarps; ++w) if (load[w] < load[best]) best = w;
        load[best] += it.cost;
        (it.kind == 0 ? mw : dw)[it.idx] = best;
    }
    double total = 0;
    for (double l : load) total += l;
    // the tensor pipe is shared by the warps of an SM sub-partition: a phase cannot beat total/4
    return std::max(*std::max_element(load, load + kMmaWarps), total / 4.0);
}

// Chooses the decomposition (m-tiles per task, max n-tiles per task) of one phase and appends its tasks, grouped by warp.
void emit_phase(MSchedule& s, MStage& st, int phase, const std::vector<FJob>& fj, const std::vector<DJob>& dj) {
    Cand best;
    std::vector<int> bmw, bdw;
    double bestspan = -1;
    const int dmt_hi = dj.empty() ? 1 : kDMT, dc_hi = dj.empty() ? 1 : kDNC, fc_hi = fj.empty() ? 1 : kNC;
    for (int cmax = fc_hi; cmax >= 1; --cmax)
        for (int dmt = dmt_hi; dmt >= 1; --dmt)
            for (int dcmax = dc_hi; dcmax >= 1; --dcmax) {
                Cand c;
                gen_ftasks(fj, s, cmax, c);
                gen_dtasks(dj, s, dmt, dcmax, c);
                std::vector<int> mw, dw;
                const double span = assign(c, mw, dw);
                if (bestspan < 0 || span < bestspan * 0.98) { bestspan = span; best = std::move(c); bmw = mw; bdw = dw; }
            }
    const bool g1 = (phase == PH_DW1G1);
    for (int w = 0; w < kMmaWarps; ++w) {
        if (phase == PH_DW3 || phase == PH_DW2 || phase == PH_DW1G1) {
            st.task_begin[phase][w] = (int)s.dtasks.size();
            for (size_t i = 0; i < best.dt.size(); ++i) if (bdw[i] == w) s.dtasks.push_back(best.dt[i]);
            st.task_begin[phase][w + 1] = (int)s.dtasks.size();
        }
        if (g1 || !(phase == PH_DW3 || phase == PH_DW2)) {
            int* tb = g1 ? st.g_begin : st.task_begin[phase];
            tb[w] = (int)s.mtasks.size();
            for (size_t i = 0; i < best.mt.size(); ++i) if (bmw[i] == w) s.mtasks.push_back(best.mt[i]);
            tb[w + 1] = (int)s.mtasks.size();
        }
    }
}

struct Group { std::vector<int> nodes; int cols = 0; };

std::string build_mschedule(const Plan& p, MmaPlan& m, const std::vector<NodeOps>& ops, MSchedule& s, bool bwd,
                            int forced_tm) {
    const int d = p.d, dc = p.dc;
    const int fixed = (bwd ? 2 : 1) * (d + dc) + 2;
    int need_min = 0;
    for (const auto& n : p.nodes) need_min = std::max(need_min, 2 * round8(n.cout) + 4 * round8(n.h));
    struct Try { int tm; int budget; int ctas; };
    const Try tries[] = {{64, (kSmemMax + 1024) / 2 - 1024, 2}, {64, kSmemMax, 1}, {32, kSmemMax, 1}};
    int TM = 0, avail = 0, raw_floats = 0;
    for (const Try& t : tries) {
        if (forced_tm > 0 && t.tm != forced_tm) continue;
        const int raw = bwd ? t.tm : kMmaThreads;   // per-sample dJ (backward) / per-thread log-det partials (forward)
        const int cols = (t.budget / 4 - raw) / (t.tm + 4);
        if (cols - fixed >= need_min) { TM = t.tm; avail = cols - fixed; s.ctas_per_sm = t.ctas; raw_floats = raw; break; }
    }
    if (TM == 0) return "hidden width too large for the warp-MMA kernel's shared-memory budget";
    s.TM = TM;

    std::vector<Group> groups;
    for (int depth = 0; depth <= p.max_depth; ++depth) {
        Group cur;
        for (size_t i = 0; i < p.nodes.size(); ++i) {
            const auto& n = p.nodes[i];
            if (n.depth != depth) continue;
            const int w = 2 * round8(n.cout) + 4 * round8(n.h);
            if (!cur.nodes.empty() && cur.cols + w > avail) { groups.push_back(cur); cur = Group(); }
            cur.nodes.push_back((int)i);
            cur.cols += w;
        }
        if (!cur.nodes.empty()) groups.push_back(cur);
    }
    int maxcols = 0;
    for (const auto& g : groups) maxcols = std::max(maxcols, g.cols);

    int col = 0;
    s.col_x = col; col += d + dc;
    if (bwd) { s.col_d = col; col += d + dc; }
    s.col_one = col++; s.col_zero = col++;
    s.col_out = col;
    s.ncols = col + maxcols;
    s.raw_off = s.ncols * (TM + 4);
    s.smem_bytes = (size_t)(s.raw_off + raw_floats) * 4;
    if (s.smem_bytes > (size_t)kSmemMax) return "internal error: warp-MMA schedule exceeds shared memory";

    const int dc8 = round8(std::max(dc, 1));
    for (const auto& g : groups) {
        MStage st{};
        int outc = 0, hc = 0;
        for (int ni : g.nodes) { outc += 2 * round8(p.nodes[ni].cout); hc += 2 * round8(p.nodes[ni].h); }
        const int col_h1 = s.col_out + outc, col_h2 = col_h1 + hc;
        int gc_off = -1;
        if (bwd && dc > 0) {   // stacked condition-gradient operand of this stage: K = stage h1 columns, N = dc8
            gc_off = (int)m.n_packed;
            m.n_packed += ((int64_t)hc * dc8 + 3) & ~int64_t(3);
            m.pack_src.resize((size_t)m.n_packed, -1);
        }
        std::vector<FJob> fj[PH_COUNT];
        std::vector<DJob> djb[PH_COUNT];
        st.ep_begin = (int)s.eps.size();
        int hoff = 0, ooff = 0;
        for (int ni : g.nodes) {
            const auto& n = p.nodes[ni];
            const NodeOps& o = ops[ni];
            const int h8 = round8(n.h), c8 = round8(n.cout), k8 = round8(std::max(n.k, 1));
            const int xin = s.col_x + n.lo, cin_col = s.col_x + d;
            fj[PH_L1].push_back(FJob{o.w1, 2 * h8 / 8, o.b1, xin, n.k, cin_col, dc, (n.cin + 7) / 8, col_h1 + hoff, 2 * h8, MT_RELU, true});
            for (int net = 0; net < 2; ++net) {
                const int h1c = col_h1 + hoff + net * h8, h2c = col_h2 + hoff + net * h8, oc = s.col_out + ooff + net * c8;
                fj[PH_L2].push_back(FJob{o.w2[net], h8 / 8, o.b2[net], h1c, h8, 0, 0, h8 / 8, h2c, h8, MT_RELU});
                fj[PH_L3].push_back(FJob{o.w3[net], c8 / 8, o.b3[net], h2c, h8, 0, 0, h8 / 8, oc, c8, 0});
                if (bwd) {
                    fj[PH_G3].push_back(FJob{o.g3[net], h8 / 8, -1, oc, c8, 0, 0, c8 / 8, h2c, h8, MT_MASK});
                    fj[PH_G2].push_back(FJob{o.g2[net], h8 / 8, -1, h2c, h8, 0, 0, h8 / 8, h1c, h8, MT_MASK});
                    djb[PH_DW3].push_back(DJob{oc, n.cout, h2c, n.h, 0, 0, o.dw[net][2], o.ld[net][2]});
                    djb[PH_DW2].push_back(DJob{h2c, n.h, h1c, n.h, 0, 0, o.dw[net][1], o.ld[net][1]});
                    djb[PH_DW1G1].push_back(DJob{h1c, n.h, xin, n.k, cin_col, dc, o.dw[net][0], o.ld[net][0]});
                    if (gc_off >= 0)
                        for (int u = 0; u < n.h; ++u)
                            for (int j = 0; j < dc; ++j)
                                m.pack_src[(size_t)gc_off + (size_t)frag(dc8 / 8, hoff + net * h8 + u, j)] =
                                    (int32_t)(poff(p, ni, net, 0, 0) + (int64_t)u * n.cin + n.k + j);
                }
            }
            if (bwd && n.k > 0)
                fj[PH_DW1G1].push_back(FJob{o.g1, k8 / 8, -1, col_h1 + hoff, 2 * h8, 0, 0, 2 * h8 / 8, s.col_d + n.lo, n.k, MT_ACCUM});
            for (int j = 0; j < n.cout; ++j)
                s.eps.push_back(Ep{n.lo + n.k + j, s.col_out + ooff + j, s.col_out + ooff + c8 + j, 0});
            hoff += 2 * h8;
            ooff += 2 * c8;
        }
        if (gc_off >= 0)
            fj[PH_DW1G1].push_back(FJob{gc_off, dc8 / 8, -1, col_h1, hc, 0, 0, hc / 8, s.col_d + d, dc, MT_ACCUM});
        st.ep_end = (int)s.eps.size();
        for (int ph = 0; ph < PH_COUNT; ++ph) {
            if (!bwd && ph > PH_L3) {
                for (int w = 0; w <= kMmaWarps; ++w) st.task_begin[ph][w] = 0;
                continue;
            }
            emit_phase(s, st, ph, fj[ph], djb[ph]);
        }
        if (!bwd) for (int w = 0; w <= kMmaWarps; ++w) st.g_begin[w] = 0;
        s.stages.push_back(st);
    }
    // linearise: one op stream per warp and direction.  A phase boundary becomes the MT_SYNC flag of the warp's next op;
    // boundaries with no op in between (the warp idles for a phase) and the final one become standalone OP_SYNC records.
    const int nprog = bwd ? 1 : 2;
    for (int pr = 0; pr < nprog; ++pr)
        for (int w = 0; w < kMmaWarps; ++w) {
            s.prog_begin[pr][w] = (int)s.prog.size();
            int pending = 0;
            auto push = [&](const void* rec) {
                WOp o;
                std::memcpy(&o, rec, sizeof(o));
                for (; pending > 1; --pending) { MTask t{}; t.type = OP_SYNC; WOp so; std::memcpy(&so, &t, sizeof(so)); s.prog.push_back(so); }
                if (pending == 1) { o.w[7] |= MT_SYNC; pending = 0; }   // flags live in the low byte of the last word of every record
                s.prog.push_back(o);
            };
            auto push_ctl = [&](int type, int a, int b) { MTask t{}; t.type = (unsigned char)type; t.w_off = a; t.b_off = b; push(&t); };
            const int ns = (int)s.stages.size();
            for (int i = 0; i < ns; ++i) {
                const MStage& st = s.stages[(bwd || pr == PROG_INV) ? i : ns - 1 - i];
                for (int ph = 0; ph < PH_COUNT; ++ph) {
                    if (!bwd && ph > PH_L3) break;
                    const bool dwp = (ph == PH_DW3 || ph == PH_DW2 || ph == PH_DW1G1);
                    if (dwp) for (int t = st.task_begin[ph][w]; t < st.task_begin[ph][w + 1]; ++t) push(&s.dtasks[t]);
                    if (ph == PH_DW1G1) for (int t = st.g_begin[w]; t < st.g_begin[w + 1]; ++t) push(&s.mtasks[t]);
                    if (!dwp) for (int t = st.task_begin[ph][w]; t < st.task_begin[ph][w + 1]; ++t) push(&s.mtasks[t]);
                    ++pending;
                    if (ph == PH_L3) { push_ctl(OP_COUPLE, st.ep_begin, st.ep_end); ++pending; }
                }
            }
            for (; pending > 0; --pending) { MTask t{}; t.type = OP_SYNC; WOp so; std::memcpy(&so, &t, sizeof(so)); s.prog.push_back(so); }
            for (int e = 0; e < 3; ++e) { MTask t{}; t.type = OP_END; WOp so; std::memcpy(&so, &t, sizeof(so)); s.prog.push_back(so); }   // look-ahead padding
            s.prog_end[pr] = (int)s.prog.size();
        }
    for (int pr = 0; pr < nprog; ++pr)
        s.fits_param[pr] = (s.prog_end[pr] - s.prog_begin[pr][0] <= kMaxProgOps) && ((int)s.eps.size() <= kMaxEps);
    s.ok = true;
    return "";
}

}  // namespace

void build_mma_plan(const Plan& p, MmaPlan& m) {
    m = MmaPlan();
    std::vector<NodeOps> ops;
    pack_nodes(p, m, ops);
    int tm_f = 0, tm_b = 0;
#ifdef HINT_B200_DEV   // tile-size overrides (developer builds only)
    if (const char* e = std::getenv("HINT_B200_MMA_TM_FWD")) tm_f = std::atoi(e);
    if (const char* e = std::getenv("HINT_B200_MMA_TM_BWD")) tm_b = std::atoi(e);
#endif
    std::string err = build_mschedule(p, m, ops, m.fwd, false, tm_f);
    if (err.empty()) err = build_mschedule(p, m, ops, m.bwd, true, tm_b);
    if (!err.empty()) { m.ok = false; m.why = err; return; }
    if (m.n_packed > (int64_t)1 << 30 || m.n_partial > (int64_t)1 << 30) { m.ok = false; m.why = "block too large"; return; }
    m.ok = true;
}

}  // namespace hint
