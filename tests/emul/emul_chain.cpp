// Host emulation of the register-chained warp-MMA kernels (chain_kernels.cuh).  TEST INFRASTRUCTURE ONLY: built and loaded
// by tests/test_emul_chain.py on CPU-only machines; never linked into libhint_b200.so, never used by the product.
// Reuses the fiber scheduler and the mma.sync model of emul_mma.cpp.
#include "emul_mma.cpp"

#include "../../hint_b200/csrc/plan_chain.h"
#include "../../hint_b200/csrc/chain_kernels.cuh"

namespace {

struct CFwdArgs {
    ChainTables T; const ChainNode* P; float* S; const float *x, *c, *W; float *z, *logdet; long long B; int bid, nblocks;
};
template <int MT, bool REV>
void cfwd_body(int tid, void* a) {
    CFwdArgs& A = *(CFwdArgs*)a;
    c_fwd_body<MT, (MT == 2 ? 8 : 16), REV, false>(A.T, A.P, A.S, A.x, A.c, A.W, A.z, A.logdet, A.B, tid, A.bid, A.nblocks);
}

struct CBwdArgs {
    ChainTables T; const ChainNode* P; ChainBwdSmem L; float* S; const float *z, *c, *W, *dz, *dl; float *x_rec, *dx, *dc, *partials;
    long long n_partial, B; int bid, nblocks;
};
template <int MT, int NW>
void cbwd_body(int tid, void* a) {
    CBwdArgs& A = *(CBwdArgs*)a;
    c_bwd_body<MT, NW>(A.T, A.P, A.L, A.S, A.z, A.c, A.W, A.dz, A.dl, A.x_rec, A.dx, A.dc, A.partials, A.n_partial, A.B, tid, A.bid, A.nblocks);
}

}  // namespace

extern "C" {
// info: {ok, n_packed, n_partial, n_nodes, max_nh}
int emul_chain_run(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits, int min_split_size,
                   const float* params, const float* x, const float* c, long long B, int rev, int nctas, int mt,
                   float* z, float* logdet, const float* dz, const float* dl, float* x_rec, float* dx, float* dcond,
                   float* dparams, long long* info) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    ChainPlan cp;
    build_chain_plan(p, cp);
    info[0] = cp.ok; info[1] = cp.n_packed; info[2] = cp.n_partial; info[3] = cp.n_nodes; info[4] = cp.max_nh; info[6] = cp.n_fwd_packed;
    if (!cp.ok) return 50;
    std::vector<float> W((size_t)cp.n_packed + 4);
    for (int64_t i = 0; i < cp.n_packed; ++i) { float lo; m_pack_elem(cp.pack_src[(size_t)i], params, W[(size_t)i], lo); }
    ChainTables T{cp.n_nodes, p.d, p.dc, p.alpha, (int)cp.n_fwd_packed, 0};
    const int bwd_cfg = mt;
    mt = mt == 2 ? 2 : 1;     // forward configuration
    {
        const int RW = 16 * mt;
        const long long ntiles = (B + RW - 1) / RW;
        const int nw = mt == 2 ? 8 : 16;
        const int nb = (int)std::max<long long>(1, std::min<long long>(nctas, (ntiles + nw - 1) / nw));
        const size_t wf = mt == 2 ? chain_fwd_warp_floats<2>(p.d, p.dc) : chain_fwd_warp_floats<1>(p.d, p.dc);
        for (int bid = 0; bid < nb; ++bid) {
            std::vector<float> S(wf * nw + 16, NAN);
            CFwdArgs A{T, cp.param.nodes, S.data(), x, c, W.data(), z, logdet, B, bid, nb};
            void (*fn)(int, void*) = mt == 2 ? (rev ? cfwd_body<2, true> : cfwd_body<2, false>) : (rev ? cfwd_body<1, true> : cfwd_body<1, false>);
            emu::run_cta(32 * nw, fn, &A);
        }
    }
    if (dz) {
        // bwd configurations: mt 1 -> (1, 4); mt 2 -> (2, 4); mt 3 -> (1, 8)
        const int bmt = bwd_cfg == 2 ? 2 : 1, bnw = bwd_cfg == 3 ? 8 : 4, TM = 16 * bmt * bnw;
        const ChainBwdSmem L = bmt == 2 ? chain_bwd_smem<2, 4>(p.d, p.dc, cp.max_nh, cp.max_no, cp.n_nodes)
                               : (bnw == 8 ? chain_bwd_smem<1, 8>(p.d, p.dc, cp.max_nh, cp.max_no, cp.n_nodes)
                                           : chain_bwd_smem<1, 4>(p.d, p.dc, cp.max_nh, cp.max_no, cp.n_nodes));
        info[5] = L.total * 4;
        const long long ntiles = (B + TM - 1) / TM;
        const int nb = (int)std::max<long long>(1, std::min<long long>(nctas, ntiles));
        std::vector<float> partials((size_t)nb * cp.n_partial, NAN);
        for (int bid = 0; bid < nb; ++bid) {
            std::vector<float> S((size_t)L.total + 16, NAN);
            std::memcpy(S.data() + L.nodes, cp.param.nodes, sizeof(ChainNode) * cp.n_nodes);
            CBwdArgs A{T, reinterpret_cast<const ChainNode*>(S.data() + L.nodes), L, S.data(), z, c, W.data(), dz, dl, x_rec, dx, dcond,
                       partials.data(), cp.n_partial, B, bid, nb};
            void (*fn)(int, void*) = bmt == 2 ? cbwd_body<2, 4> : (bnw == 8 ? cbwd_body<1, 8> : cbwd_body<1, 4>);
            emu::run_cta(32 * bnw, fn, &A);
        }
        for (int64_t i = 0; i < p.n_params; ++i) {
            float a = 0.f;
            for (int q = 0; q < nb; ++q) a += partials[(size_t)q * cp.n_partial + cp.unpack_src[(size_t)i]];
            dparams[i] = a;
        }
    }
    return 0;
}
}
