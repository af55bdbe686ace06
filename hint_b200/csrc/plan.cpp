// Plan construction for one HINT coupling block (host only, no CUDA).
//
// Tree rules restated from the reference, /root/reference/hint.py:
//   widths   : c_internal empty -> [d]; consumed one per level, last entry repeats   (hint.py:29-34,50-52)
//   split    : upper = first floor(w/2) columns, lower = the rest                    (hint.py:41,68)
//   subnets  : s,t : (k + dc) -> h -> h -> (w - k)                                    (hint.py:44-45)
//   internal : w >= 2*min_split_size and max_splits != 0; children get max_splits-1   (hint.py:47-52)
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <functional>

namespace hint {

namespace {

int64_t& poff(Plan& p, int node, int net, int layer, int kind) {
    return p.param_offsets[(size_t)node * 12 + net * 6 + layer * 2 + kind];
}

void layer_dims(const hint_node_info_t& n, int layer, int& K, int& N) {
    if (layer == 0) { K = n.cin; N = n.h; }
    else if (layer == 1) { K = n.h; N = n.h; }
    else { K = n.h; N = n.cout; }
}

void build_tree(Plan& p) {
    std::function<int(int, int, int, int, int)> rec = [&](int lo, int hi, int depth, int splits_left, int parent) {
        const int w = hi - lo;
        hint_node_info_t n{};
        n.depth = depth; n.lo = lo; n.hi = hi; n.k = w / 2;
        n.cin = n.k + p.dc;
        n.h = p.widths[std::min<size_t>(depth, p.widths.size() - 1)];
        n.cout = w - n.k;
        const bool internal = (w >= 2 * p.min_split_size) && (splits_left != 0);
        n.leaf = internal ? 0 : 1;
        n.parent = parent; n.upper = -1; n.lower = -1; n.param_offset = 0;
        const int idx = (int)p.nodes.size();
        p.nodes.push_back(n);
        p.max_depth = std::max(p.max_depth, depth);
        if (internal) {
            const int u = rec(lo, lo + n.k, depth + 1, splits_left - 1, idx);
            const int l = rec(lo + n.k, hi, depth + 1, splits_left - 1, idx);
            p.nodes[idx].upper = u;
            p.nodes[idx].lower = l;
        }
        return idx;
    };
    rec(0, p.d, 0, p.max_splits, -1);
}

void build_canonical_layout(Plan& p) {
    p.param_offsets.assign(p.nodes.size() * 12, 0);
    int64_t off = 0, flops = 0;
    for (size_t i = 0; i < p.nodes.size(); ++i) {
        auto& n = p.nodes[i];
        n.param_offset = off;
        for (int net = 0; net < 2; ++net)
            for (int layer = 0; layer < 3; ++layer) {
                int K, N;
                layer_dims(n, layer, K, N);
                poff(p, (int)i, net, layer, 0) = off; off += (int64_t)K * N;
                poff(p, (int)i, net, layer, 1) = off; off += N;
                flops += 2LL * K * N;
            }
    }
    p.n_params = off;
    p.flops = flops;
}

// Packed weights: forward operands transposed to [K][N4], backward operands in [K'][N4'] form, all
// 4-float aligned so every weight fetch is one 128-bit load.
void build_packed_layout(Plan& p) {
    p.packs.assign(p.nodes.size(), NodePack{});
    int64_t off = 0, poffp = 0;
    auto alloc = [&](int64_t n) { int64_t o = off; off += (n + 3) & ~int64_t(3); return (int)o; };
    for (size_t i = 0; i < p.nodes.size(); ++i) {
        const auto& n = p.nodes[i];
        NodePack& pk = p.packs[i];
        const int hp = round4(n.h), kp = round4(n.k);
        for (int net = 0; net < 2; ++net) {
            for (int layer = 0; layer < 3; ++layer) {
                int K, N;
                layer_dims(n, layer, K, N);
                pk.wt[net][layer] = alloc((int64_t)K * round4(N));
                pk.b[net][layer] = alloc(round4(N));
                pk.dw[net][layer] = (int)poffp;
                poffp += (int64_t)round4(N) * round16(K);
                pk.db[net][layer] = (int)poffp;
                poffp += round4(N);
            }
            pk.wc3[net] = alloc((int64_t)n.cout * hp);
            pk.wc2[net] = alloc((int64_t)n.h * hp);
        }
        pk.wg1 = alloc((int64_t)2 * hp * kp);
    }
    p.n_packed = off;
    p.n_partial = poffp;
    p.pack_src.assign((size_t)off, -1);
    p.unpack_src.assign((size_t)p.n_params, -1);
    for (size_t i = 0; i < p.nodes.size(); ++i) {
        const auto& n = p.nodes[i];
        const NodePack& pk = p.packs[i];
        const int hp = round4(n.h), kp = round4(n.k);
        for (int net = 0; net < 2; ++net) {
            for (int layer = 0; layer < 3; ++layer) {
                int K, N;
                layer_dims(n, layer, K, N);
                const int Np = round4(N), ld = round16(K);
                const int64_t w = poff(p, (int)i, net, layer, 0), b = poff(p, (int)i, net, layer, 1);
                for (int nn = 0; nn < N; ++nn) {
                    for (int k = 0; k < K; ++k) {
                        p.pack_src[(size_t)pk.wt[net][layer] + (size_t)k * Np + nn] = (int32_t)(w + (int64_t)nn * K + k);
                        p.unpack_src[(size_t)(w + (int64_t)nn * K + k)] = pk.dw[net][layer] + nn * ld + k;
                    }
                    p.pack_src[(size_t)pk.b[net][layer] + nn] = (int32_t)(b + nn);
                    p.unpack_src[(size_t)(b + nn)] = pk.db[net][layer] + nn;
                }
            }
            const int64_t w3 = poff(p, (int)i, net, 2, 0), w2 = poff(p, (int)i, net, 1, 0), w1 = poff(p, (int)i, net, 0, 0);
            for (int r = 0; r < n.cout; ++r)
                for (int u = 0; u < n.h; ++u) p.pack_src[(size_t)pk.wc3[net] + (size_t)r * hp + u] = (int32_t)(w3 + (int64_t)r * n.h + u);
            for (int r = 0; r < n.h; ++r)
                for (int u = 0; u < n.h; ++u) p.pack_src[(size_t)pk.wc2[net] + (size_t)r * hp + u] = (int32_t)(w2 + (int64_t)r * n.h + u);
            for (int u = 0; u < n.h; ++u)
                for (int j = 0; j < n.k; ++j)
                    p.pack_src[(size_t)pk.wg1 + (size_t)(net * hp + u) * kp + j] = (int32_t)(w1 + (int64_t)u * n.cin + j);
        }
    }
}

struct StageNodes { std::vector<int> nodes; int h1_cols = 0, out_cols = 0; };

std::string build_schedule(Plan& p, Schedule& s, bool bwd, int forced_tm) {
    const int d = p.d, dc = p.dc;
    s.DX = round4(d + dc);
    int cap_min = 0;
    std::vector<int> level_out(p.max_depth + 1, 0);
    for (const auto& n : p.nodes) {
        cap_min = std::max(cap_min, 2 * round4(n.h));
        level_out[n.depth] += 2 * round4(n.cout);
    }
    const int out_ub = *std::max_element(level_out.begin(), level_out.end());
    const int fixed = bwd ? (2 * s.DX + 4) : s.DX;
    const int raw_floats = kThreads;  // log-det partials (fwd) or dJ per sample (bwd)

    static const int kTms[] = {128, 64, 32, 16, 8};
    int TM = 0, cap_limit = 0;
    for (int tm : kTms) {
        if (forced_tm > 0 && tm != forced_tm) continue;
        const int avail = (kSmemMax - raw_floats * 4) / ((tm + 4) * 4);
        const int cap = ((avail - fixed - out_ub) / 2) & ~3;
        if (cap >= cap_min) { TM = tm; cap_limit = cap; break; }
    }
    if (TM == 0) return "hidden width too large for the fused kernel's shared-memory budget";
    s.TM = TM;

    // greedy grouping of each level's nodes (pre-order) into stages under the column capacity
    std::vector<StageNodes> groups;
    for (int depth = 0; depth <= p.max_depth; ++depth) {
        StageNodes cur;
        for (size_t i = 0; i < p.nodes.size(); ++i) {
            const auto& n = p.nodes[i];
            if (n.depth != depth) continue;
            const int w = 2 * round4(n.h);
            if (!cur.nodes.empty() && cur.h1_cols + w > cap_limit) { groups.push_back(cur); cur = StageNodes(); }
            cur.nodes.push_back((int)i);
            cur.h1_cols += w;
            cur.out_cols += 2 * round4(n.cout);
        }
        if (!cur.nodes.empty()) groups.push_back(cur);
    }
    int cap = 0, outc = 0;
    for (const auto& g : groups) { cap = std::max(cap, g.h1_cols); outc = std::max(outc, g.out_cols); }

    int col = 0;
    s.col_x = col; col += s.DX;
    if (bwd) {
        s.col_d = col; col += s.DX;
        s.col_one = col; s.col_zero = col + 1; col += 4;
    }
    s.col_out = col; col += outc;
    s.col_h1 = col; col += cap;
    s.col_h2 = col; col += cap;
    s.ncols = col;
    s.raw_off = col * (TM + 4);
    s.smem_bytes = (size_t)(s.raw_off + raw_floats) * 4;
    if (s.smem_bytes > (size_t)kSmemMax) return "internal error: schedule exceeds shared memory";

    const int dcp = round4(dc);
    for (const auto& g : groups) {
        Stage st{};
        std::vector<CG> ph[7];
        std::vector<DwJob> dw[3];
        st.ep_begin = (int)s.eps.size();
        int hoff = 0, ooff = 0;
        int wgc = -1;
        if (bwd && dc > 0) {  // stacked condition-gradient operand of this stage: [stage H1 cols][dcp]
            wgc = (int)p.n_packed;
            p.n_packed += (int64_t)g.h1_cols * dcp;
            p.pack_src.resize((size_t)p.n_packed, -1);
        }
        for (int ni : g.nodes) {
            const auto& n = p.nodes[ni];
            const NodePack& pk = p.packs[ni];
            const int hp = round4(n.h), kp = round4(n.k), cp = round4(n.cout);
            for (int net = 0; net < 2; ++net) {
                const int h1c = s.col_h1 + hoff + net * hp, h2c = s.col_h2 + hoff + net * hp;
                const int oc = s.col_out + ooff + net * cp;
                for (int n0 = 0; n0 < hp; n0 += 4) {
                    ph[0].push_back(CG{pk.wt[net][0] + n0, hp, pk.b[net][0] + n0, s.col_x + n.lo, n.k, s.col_x + d, dc, h1c + n0, 4, CG_RELU, 0, 0});
                    ph[1].push_back(CG{pk.wt[net][1] + n0, hp, pk.b[net][1] + n0, h1c, n.h, 0, 0, h2c + n0, 4, CG_RELU, 0, 0});
                    if (bwd) {
                        ph[3].push_back(CG{pk.wc3[net] + n0, hp, -1, oc, n.cout, 0, 0, h2c + n0, 4, CG_MASK, 0, 0});
                        ph[4].push_back(CG{pk.wc2[net] + n0, hp, -1, h2c, n.h, 0, 0, h1c + n0, 4, CG_MASK, 0, 0});
                    }
                }
                for (int n0 = 0; n0 < cp; n0 += 4)
                    ph[2].push_back(CG{pk.wt[net][2] + n0, cp, pk.b[net][2] + n0, h2c, n.h, 0, 0, oc + n0, 4, 0, 0, 0});
                if (bwd) {
                    dw[0].push_back(DwJob{0, cp / 4, round16(n.h) / 16, oc, h2c, n.h, 0, 0, pk.dw[net][2], round16(n.h), pk.db[net][2], 0});
                    dw[1].push_back(DwJob{0, hp / 4, round16(n.h) / 16, h2c, h1c, n.h, 0, 0, pk.dw[net][1], round16(n.h), pk.db[net][1], 0});
                    dw[2].push_back(DwJob{0, hp / 4, round16(n.cin) / 16, h1c, s.col_x + n.lo, n.k, s.col_x + d, dc, pk.dw[net][0], round16(n.cin), pk.db[net][0], 0});
                    if (wgc >= 0)
                        for (int u = 0; u < n.h; ++u)
                            for (int j = 0; j < dc; ++j)
                                p.pack_src[(size_t)wgc + (size_t)(hoff + net * hp + u) * dcp + j] =
                                    (int32_t)(poff(p, ni, net, 0, 0) + (int64_t)u * n.cin + n.k + j);
                }
            }
            if (bwd)
                for (int n0 = 0; n0 < kp; n0 += 4)
                    ph[5].push_back(CG{pk.wg1 + n0, kp, -1, s.col_h1 + hoff, 2 * hp, 0, 0, s.col_d + n.lo + n0, std::min(4, n.k - n0), CG_ACCUM, 0, 0});
            for (int j = 0; j < n.cout; ++j)
                s.eps.push_back(Ep{n.lo + n.k + j, s.col_out + ooff + j, s.col_out + ooff + cp + j, 0});
            hoff += 2 * hp;
            ooff += 2 * cp;
        }
        if (wgc >= 0)
            for (int n0 = 0; n0 < dcp; n0 += 4)
                ph[6].push_back(CG{wgc + n0, dcp, -1, s.col_h1, g.h1_cols, 0, 0, s.col_d + d + n0, std::min(4, dc - n0), CG_ACCUM, 0, 0});
        st.ep_end = (int)s.eps.size();
        for (int q = 0; q < 7; ++q) {
            st.cg_begin[q] = (int)s.cgs.size();
            s.cgs.insert(s.cgs.end(), ph[q].begin(), ph[q].end());
        }
        st.cg_begin[7] = (int)s.cgs.size();
        for (int q = 0; q < 3; ++q) {
            st.dw_begin[q] = (int)s.dwjobs.size();
            int items = 0, bitems = 0;
            for (auto& j : dw[q]) {
                j.item_begin = items;
                j.bitem_begin = bitems;
                items += j.nN * j.nKB * 4;
                bitems += j.nN * 4;
                s.dwjobs.push_back(j);
            }
            st.dw_items[q] = items;
            st.dw_bitems[q] = bitems;
        }
        st.dw_begin[3] = (int)s.dwjobs.size();
        s.stages.push_back(st);
    }
    return "";
}

}  // namespace

std::string build_plan(Plan& p, int d, int dc, const int32_t* c_internal, int n_internal, double clamp,
                       int max_splits, int min_split_size, int reshuffle, int* code) {
    *code = HINT_ERR_INVALID;
    if (d < 2) return "dims_in[0][0] must be >= 2";
    if (dc < 0) return "negative condition width";
    if (min_split_size < 1) return "min_split_size must be >= 1";
    if (n_internal < 0 || (n_internal > 0 && c_internal == nullptr)) return "bad c_internal";
    for (int i = 0; i < n_internal; ++i)
        if (c_internal[i] < 1) return "c_internal entries must be >= 1";
    if (d + dc > 4096) return "d + dc too large for the fused kernel";
    if (reshuffle) {
        *code = HINT_ERR_UNSUPPORTED;
        return "reshuffle=True needs FrEIA's HouseholderPerm (absent from the reference tree; parity unpinned)";
    }
    p = Plan();
    p.d = d; p.dc = dc; p.clamp = clamp;
    p.alpha = (float)(clamp * 0.636);
    p.max_splits = max_splits; p.min_split_size = min_split_size;
    if (n_internal == 0) p.widths = {d};
    else p.widths.assign(c_internal, c_internal + n_internal);
    build_tree(p);
    build_canonical_layout(p);
    if (p.n_params > (int64_t)1 << 30) return "parameter count too large";
    build_packed_layout(p);
    int tm_f = 0, tm_b = 0;
#ifdef HINT_B200_DEV   // tile-size overrides (developer builds only)
    if (const char* e = std::getenv("HINT_B200_TM_FWD")) tm_f = std::atoi(e);
    if (const char* e = std::getenv("HINT_B200_TM_BWD")) tm_b = std::atoi(e);
#endif
    *code = HINT_ERR_UNSUPPORTED;
    std::string err = build_schedule(p, p.fwd, false, tm_f);
    if (!err.empty()) return err;
    err = build_schedule(p, p.bwd, true, tm_b);
    if (!err.empty()) return err;
    *code = HINT_OK;
    return "";
}

}  // namespace hint
