// Planner of the register-chained warp-MMA kernels (host only, no CUDA): groups tiny nodes into super nodes, picks the
// instantiated shape of every (super) node, packs the operands in B-fragment order for that shape and lays out the
// partial-gradient buffer.  Tree rules come from plan.cpp (hint.py:25-54); per-node math from hint.py:62-101.
#include "plan_chain.h"

#include <algorithm>
#include <cstring>

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
// block-diagonal operand: fragment b holds B[8b + kk][8b + nn]
inline int64_t frag_bd(int b, int kk, int nn) { return ((int64_t)b * 32 + nn * 4 + (kk >> 1)) * 2 + (kk & 1); }

// float index of C[m][n] inside a C-fragment-ordered matrix of NT n-tiles (m-tile major): lane (g = m%8, t = (n%8)/2),
// element 2*(m%16 >= 8) + n%2
inline int64_t cfrag(int NT, int m, int n) {
    const int i = m >> 4, mm = m & 15, j = n >> 3, nn = n & 7;
    return ((int64_t)(i * NT + j) * 32 + (mm & 7) * 4 + (nn >> 1)) * 4 + 2 * (mm >> 3) + (nn & 1);
}

int pick_shape(int ks1, int nh, int no, int bd) {
    int best = -1;
    long long best_cost = 0;
    for (int s = 0; s < kChainNumShapes; ++s) {
        const ChainShape& c = kChainShapes[s];
        if (c.bd != bd || c.ks1 < ks1 || c.nh < nh || c.no < no) continue;
        const long long cost = (long long)c.ks1 * c.nh + (long long)(c.bd ? c.nh : c.nh * c.nh) + (long long)c.nh * c.no;
        if (best < 0 || cost < best_cost) { best = s; best_cost = cost; }
    }
    return best;
}

struct Group {
    std::vector<int> members;   // plan node indices, left to right
    int shape = -1;
};

}  // namespace

void build_chain_plan(const Plan& p, ChainPlan& c) {
    c = ChainPlan();
    const int n = (int)p.nodes.size();
    int max_depth = 0;
    for (const auto& nd : p.nodes) max_depth = std::max(max_depth, (int)nd.depth);
    // groups in forward order: deepest level first, nodes of a level left to right (pre-order index = left to right)
    std::vector<Group> groups;
    for (int depth = max_depth; depth >= 0; --depth) {
        std::vector<int> level;
        for (int i = 0; i < n; ++i)
            if (p.nodes[(size_t)i].depth == depth) level.push_back(i);
        std::sort(level.begin(), level.end(), [&](int a, int b) { return p.nodes[(size_t)a].lo < p.nodes[(size_t)b].lo; });
        size_t q = 0;
        while (q < level.size()) {
            const auto& nd = p.nodes[(size_t)level[q]];
            Group g;
            g.members.push_back(level[q]);
            const bool tiny = nd.h <= 8 && nd.cin <= 8 && nd.cout <= 8;
            ++q;
            if (tiny) {
                int kin = nd.k, kout = nd.cout;
                while (q < level.size() && (int)g.members.size() < 4) {
                    const auto& nx = p.nodes[(size_t)level[q]];
                    if (!(nx.h <= 8) || kin + nx.k + p.dc > 8 || kout + nx.cout > 8) break;
                    kin += nx.k; kout += nx.cout;
                    g.members.push_back(level[q]);
                    ++q;
                }
            }
            if (g.members.size() > 1) g.shape = pick_shape(1, (int)g.members.size(), 1, 1);
            else g.shape = pick_shape(round8(nd.cin) / 8, round8(nd.h) / 8, round8(nd.cout) / 8, 0);
            if (g.shape < 0) {
                c.why = "node " + std::to_string(g.members[0]) + " (cin " + std::to_string(nd.cin) + ", h " + std::to_string(nd.h) +
                        ", cout " + std::to_string(nd.cout) + ") exceeds the largest instantiated chain shape";
                return;
            }
            groups.push_back(g);
        }
    }
    const int ng = (int)groups.size();
    if (ng > kChainMaxNodes) { c.why = "tree has more than " + std::to_string(kChainMaxNodes) + " (super) nodes"; return; }
    std::vector<int64_t> w_off(ng), wt_off(ng), dw_off(ng);
    int64_t off = 0, toff = 0, doff = 0;
    for (int q = 0; q < ng; ++q) {
        const ChainShape& cs = kChainShapes[groups[q].shape];
        const ChainOffRt o = chain_off(cs);
        w_off[q] = off; off += 2 * (int64_t)o.net;
        wt_off[q] = toff; toff += 2 * (int64_t)o.tnet;
        dw_off[q] = doff; doff += 2 * (int64_t)o.dnet;
        // buffer slots (n-tiles): nodes with nh <= 4 keep both nets' activations at once (chain_kernels.cuh: c_node_bwd)
        c.max_nh = std::max(c.max_nh, cs.nh <= 4 ? 2 * cs.nh : cs.nh);
        c.max_no = std::max(c.max_no, cs.nh <= 4 ? 2 * cs.no : cs.no);
    }
    c.n_fwd_packed = off;
    c.n_packed = off + toff;
    c.n_partial = doff;
    c.pack_src.assign((size_t)c.n_packed, -1);
    c.unpack_src.assign((size_t)p.n_params, -1);
    c.n_nodes = ng;
    std::memset(&c.param, 0, sizeof(c.param));
    for (int q = 0; q < ng; ++q) {
        const Group& g = groups[q];
        const ChainShape& cs = kChainShapes[g.shape];
        const ChainOffRt o = chain_off(cs);
        const int NH = cs.nh, NO = cs.no, KS1 = cs.ks1;
        ChainNode& cn = c.param.nodes[q];
        cn.shape = g.shape;
        cn.w_off = (int)w_off[q]; cn.wt_off = (int)(off + wt_off[q]); cn.dw_off = (int)dw_off[q];
        for (int f = 0; f < kChainMaxIn; ++f) cn.in_col[f] = -1;
        for (int f = 0; f < kChainMaxOut; ++f) cn.out_col[f] = -1;
        // feature / output layout of the group: own upper columns of the members back to back, then the shared condition
        int ksum = 0;
        for (int m : g.members) ksum += p.nodes[(size_t)m].k;
        const int fc = ksum;     // first condition feature
        for (int j = 0; j < p.dc; ++j) cn.in_col[fc + j] = (short)(p.d + j);
        int fo = 0, co = 0;
        for (size_t mi = 0; mi < g.members.size(); ++mi) {
            const int i = g.members[mi];
            const auto& nd = p.nodes[(size_t)i];
            const int hb = cs.bd ? 8 * (int)mi : 0;      // first hidden unit of this member
            for (int f = 0; f < nd.k; ++f) cn.in_col[fo + f] = (short)(nd.lo + f);
            for (int r = 0; r < nd.cout; ++r) cn.out_col[co + r] = (short)(nd.lo + nd.k + r);
            auto feat = [&](int f) { return f < nd.k ? fo + f : fc + (f - nd.k); };     // subnet input f -> group feature
            for (int net = 0; net < 2; ++net) {
                const int64_t base = w_off[q] + (int64_t)net * o.net;
                const int64_t tbase = off + wt_off[q] + (int64_t)net * o.tnet;
                const int64_t dbase = dw_off[q] + (int64_t)net * o.dnet;
                const int64_t w1 = poff(p, i, net, 0, 0), b1 = poff(p, i, net, 0, 1);
                const int64_t w2 = poff(p, i, net, 1, 0), b2 = poff(p, i, net, 1, 1);
                const int64_t w3 = poff(p, i, net, 2, 0), b3 = poff(p, i, net, 2, 1);
                auto set = [&](int64_t idx, int64_t src) { c.pack_src[(size_t)(base + idx)] = (int32_t)src; };
                auto sett = [&](int64_t idx, int64_t src) { c.pack_src[(size_t)(tbase + idx)] = (int32_t)src; };
                auto set_bias = [&](int64_t reg, int u, int64_t src) { c.pack_src[(size_t)(base + reg + u)] = (int32_t)(-src - 2); };
                auto un = [&](int64_t param, int64_t idx) { c.unpack_src[(size_t)param] = (int32_t)(dbase + idx); };
                for (int u = 0; u < nd.h; ++u) {
                    for (int f = 0; f < nd.cin; ++f) {
                        const int64_t src = w1 + (int64_t)u * nd.cin + f;
                        set(o.w1 + frag(NH, feat(f), hb + u), src);                 // B[f][u] = W1[u][f]
                        sett(o.w1t + frag(KS1, hb + u, feat(f)), src);              // B[u][f] = W1[u][f]
                        un(src, o.dw1 + cfrag(NH, feat(f), hb + u));
                    }
                    set_bias(o.b1, hb + u, b1 + u);
                    un(b1 + u, o.dw1 + cfrag(NH, 8 * KS1, hb + u));
                    for (int v = 0; v < nd.h; ++v) {
                        const int64_t src = w2 + (int64_t)u * nd.h + v;
                        if (cs.bd) {
                            set(o.w2 + frag_bd((int)mi, v, u), src);                // B[v][u] = W2[u][v]
                            sett(o.w2t + frag_bd((int)mi, u, v), src);              // B[u][v] = W2[u][v]
                        } else {
                            set(o.w2 + frag(NH, v, u), src);
                            sett(o.w2t + frag(NH, u, v), src);
                        }
                        un(src, o.dw2 + cfrag(NH, hb + v, hb + u));
                    }
                    set_bias(o.b2, hb + u, b2 + u);
                    un(b2 + u, o.dw2 + cfrag(NH, 8 * NH, hb + u));
                }
                for (int r = 0; r < nd.cout; ++r) {
                    for (int v = 0; v < nd.h; ++v) {
                        const int64_t src = w3 + (int64_t)r * nd.h + v;
                        set(o.w3 + frag(NO, hb + v, co + r), src);                  // B[v][r] = W3[r][v]
                        sett(o.w3t + frag(NH, co + r, hb + v), src);                // B[r][v] = W3[r][v]
                        un(src, o.dw3 + cfrag(NO, hb + v, co + r));
                    }
                    set_bias(o.b3, co + r, b3 + r);
                    un(b3 + r, o.dw3 + cfrag(NO, 8 * NH, co + r));
                }
            }
            fo += nd.k;
            co += nd.cout;
        }
    }
    c.ok = true;
}

}  // namespace hint
