// Planner of the register-chained warp-MMA kernels (host only, no CUDA): picks the instantiated shape of every node,
// packs the operands in B-fragment order for that shape and lays out the partial-gradient buffer.
// Tree rules come from plan.cpp (hint.py:25-54); per-node math from hint.py:62-101 (see chain_kernels.cuh).
#include "plan_chain.h"

#include <algorithm>
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

// float index of C[m][n] inside a C-fragment-ordered matrix of NT n-tiles (m-tile major): lane (g = m%8, t = (n%8)/2),
// element 2*(m%16 >= 8) + n%2
inline int64_t cfrag(int NT, int m, int n) {
    const int i = m >> 4, mm = m & 15, j = n >> 3, nn = n & 7;
    return ((int64_t)(i * NT + j) * 32 + (mm & 7) * 4 + (nn >> 1)) * 4 + 2 * (mm >> 3) + (nn & 1);
}

int pick_shape(int ks1, int nh, int no) {
    int best = -1;
    long long best_cost = 0;
    for (int s = 0; s < kChainNumShapes; ++s) {
        const ChainShape& c = kChainShapes[s];
        if (c.ks1 < ks1 || c.nh < nh || c.no < no) continue;
        const long long cost = (long long)c.ks1 * c.nh + (long long)c.nh * c.nh + (long long)c.nh * c.no;
        if (best < 0 || cost < best_cost) { best = s; best_cost = cost; }
    }
    return best;
}

}  // namespace

void build_chain_plan(const Plan& p, ChainPlan& c) {
    c = ChainPlan();
    const int n = (int)p.nodes.size();
    if (n > kChainMaxNodes) { c.why = "tree has more than " + std::to_string(kChainMaxNodes) + " nodes"; return; }
    // forward order: children before their parent
    std::vector<int> order;
    std::function<void(int)> visit = [&](int i) {
        const auto& nd = p.nodes[(size_t)i];
        if (!nd.leaf) { visit(nd.upper); visit(nd.lower); }
        order.push_back(i);
    };
    visit(0);
    std::vector<int> shape(n), w_off(n), wt_off(n), dw_off(n);
    int64_t off = 0, toff = 0, doff = 0;
    for (int i = 0; i < n; ++i) {
        const auto& nd = p.nodes[(size_t)i];
        const int s = pick_shape(round8(nd.cin) / 8, round8(nd.h) / 8, round8(nd.cout) / 8);
        if (s < 0) {
            c.why = "node " + std::to_string(i) + " (cin " + std::to_string(nd.cin) + ", h " + std::to_string(nd.h) + ", cout " +
                    std::to_string(nd.cout) + ") exceeds the largest instantiated chain shape";
            return;
        }
        shape[i] = s;
        const ChainShape& cs = kChainShapes[s];
        w_off[i] = (int)off;
        off += 2 * (int64_t)chain_net_floats(cs.ks1, cs.nh, cs.no);
        wt_off[i] = (int)toff;
        toff += 2 * (int64_t)chain_tnet_floats(cs.ks1, cs.nh, cs.no);
        dw_off[i] = (int)doff;
        doff += 2 * (int64_t)chain_dw_net_floats(cs.ks1, cs.nh, cs.no);
        c.max_nh = std::max(c.max_nh, cs.nh);
        c.max_no = std::max(c.max_no, cs.no);
    }
    c.n_fwd_packed = off;
    c.n_packed = off + toff;
    for (int i = 0; i < n; ++i) wt_off[i] += (int)off;
    c.n_partial = doff;
    c.pack_src.assign((size_t)c.n_packed, -1);
    c.unpack_src.assign((size_t)p.n_params, -1);
    for (int i = 0; i < n; ++i) {
        const auto& nd = p.nodes[(size_t)i];
        const ChainShape& cs = kChainShapes[shape[i]];
        const int KS1 = cs.ks1, NH = cs.nh, NO = cs.no;
        for (int net = 0; net < 2; ++net) {
            const int64_t base = w_off[i] + (int64_t)net * chain_net_floats(KS1, NH, NO);
            const int64_t tbase = wt_off[i] + (int64_t)net * chain_tnet_floats(KS1, NH, NO);
            const int64_t w1 = poff(p, i, net, 0, 0), b1 = poff(p, i, net, 0, 1);
            const int64_t w2 = poff(p, i, net, 1, 0), b2 = poff(p, i, net, 1, 1);
            const int64_t w3 = poff(p, i, net, 2, 0), b3 = poff(p, i, net, 2, 1);
            auto set = [&](int64_t idx, int64_t src) { c.pack_src[(size_t)(base + idx)] = (int32_t)src; };
            auto sett = [&](int64_t idx, int64_t src) { c.pack_src[(size_t)(tbase + idx)] = (int32_t)src; };
            auto set_exact = [&](int64_t idx, int64_t src) { c.pack_src[(size_t)(base + idx)] = (int32_t)(-src - 2); };
            for (int u = 0; u < nd.h; ++u) {
                for (int f = 0; f < nd.cin; ++f) {
                    set(chain_w1(KS1, NH, NO) + frag(NH, f, u), w1 + (int64_t)u * nd.cin + f);      // B[f][u] = W1[u][f]
                    sett(chain_w1t(KS1, NH, NO) + frag(KS1, u, f), w1 + (int64_t)u * nd.cin + f);    // B[u][f] = W1[u][f]
                }
                set_exact(chain_b1(KS1, NH, NO) + u, b1 + u);
                for (int v = 0; v < nd.h; ++v) {
                    set(chain_w2(KS1, NH, NO) + frag(NH, v, u), w2 + (int64_t)u * nd.h + v);        // B[v][u] = W2[u][v]
                    sett(chain_w2t(KS1, NH, NO) + frag(NH, u, v), w2 + (int64_t)u * nd.h + v);       // B[u][v] = W2[u][v]
                }
                set_exact(chain_b2(KS1, NH, NO) + u, b2 + u);
            }
            for (int r = 0; r < nd.cout; ++r) {
                for (int v = 0; v < nd.h; ++v) {
                    set(chain_w3(KS1, NH, NO) + frag(NO, v, r), w3 + (int64_t)r * nd.h + v);        // B[v][r] = W3[r][v]
                    sett(chain_w3t(KS1, NH, NO) + frag(NH, r, v), w3 + (int64_t)r * nd.h + v);       // B[r][v] = W3[r][v]
                }
                set_exact(chain_b3(KS1, NH, NO) + r, b3 + r);
            }
            // partial gradients: dW_lT = [In | 1]^T dOut as C fragments (rows = in features, bias row 8*k_tiles8)
            const int64_t dbase = dw_off[i] + (int64_t)net * chain_dw_net_floats(KS1, NH, NO);
            auto un = [&](int64_t param, int64_t idx) { c.unpack_src[(size_t)param] = (int32_t)(dbase + idx); };
            for (int u = 0; u < nd.h; ++u) {
                for (int f = 0; f < nd.cin; ++f) un(w1 + (int64_t)u * nd.cin + f, chain_dw1(KS1, NH, NO) + cfrag(NH, f, u));
                un(b1 + u, chain_dw1(KS1, NH, NO) + cfrag(NH, 8 * KS1, u));
                for (int v = 0; v < nd.h; ++v) un(w2 + (int64_t)u * nd.h + v, chain_dw2(KS1, NH, NO) + cfrag(NH, v, u));
                un(b2 + u, chain_dw2(KS1, NH, NO) + cfrag(NH, 8 * NH, u));
            }
            for (int r = 0; r < nd.cout; ++r) {
                for (int v = 0; v < nd.h; ++v) un(w3 + (int64_t)r * nd.h + v, chain_dw3(KS1, NH, NO) + cfrag(NO, v, r));
                un(b3 + r, chain_dw3(KS1, NH, NO) + cfrag(NO, 8 * NH, r));
            }
        }
    }
    c.n_nodes = n;
    std::memset(&c.param, 0, sizeof(c.param));
    for (int q = 0; q < n; ++q) {
        const int i = order[(size_t)q];
        const auto& nd = p.nodes[(size_t)i];
        ChainNode& cn = c.param.nodes[q];
        cn.shape = shape[i];
        cn.lo = nd.lo; cn.k = nd.k; cn.cout = nd.cout; cn.cin = nd.cin;
        cn.w_off = w_off[i]; cn.wt_off = wt_off[i]; cn.dw_off = dw_off[i];
    }
    c.ok = true;
}

}  // namespace hint
