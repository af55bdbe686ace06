// Host emulation of the FP32 fused-tree kernels.  TEST INFRASTRUCTURE ONLY (built and loaded by
// tests/test_emul.py on CPU-only machines; never linked into libhint_b200.so, never used by the
// product).  It runs the very same phase functions (simt_phases.cuh / simt_kernels.cuh) with the CTA
// replaced by a loop over thread ids, so plan/schedule/index bugs show up without a GPU.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

using std::min;
#include "../../hint_b200/csrc/plan.h"
#include "../../hint_b200/csrc/simt_kernels.cuh"

using namespace hint;

namespace {
DevTables tables(const Plan& p, const Schedule& s) {
    DevTables t;
    t.cgs = s.cgs.data(); t.eps = s.eps.data(); t.dwjobs = s.dwjobs.data(); t.stages = s.stages.data();
    t.nstages = (int)s.stages.size();
    t.d = p.d; t.dc = p.dc;
    t.col_x = s.col_x; t.col_d = s.col_d; t.col_one = s.col_one; t.col_zero = s.col_zero;
    t.raw_off = s.raw_off; t.alpha = p.alpha;
    return t;
}
std::vector<float> pack(const Plan& p, const float* params) {
    std::vector<float> w((size_t)p.n_packed + 4);
    for (int64_t i = 0; i < p.n_packed; ++i) w[i] = p.pack_src[i] < 0 ? 0.f : params[p.pack_src[i]];
    return w;
}
template <int TM>
void fwd_all(const Plan& p, const float* x, const float* c, const float* W, float* z, float* logdet, long long B, int rev) {
    std::vector<float> S(p.fwd.smem_bytes / 4 + 16, NAN);
    DevTables T = tables(p, p.fwd);
    for (long long row0 = 0; row0 < B; row0 += TM) fwd_tile<TM>(T, S.data(), x, c, W, z, logdet, B, rev, row0);
}
template <int TM>
void bwd_all(const Plan& p, int nctas, const float* z, const float* c, const float* W, const float* dz, const float* dl,
             float* x_rec, float* dx, float* dc, float* dparams, long long B) {
    std::vector<float> S(p.bwd.smem_bytes / 4 + 16, NAN);
    DevTables T = tables(p, p.bwd);
    const long long ntiles = (B + TM - 1) / TM;
    nctas = (int)std::min<long long>(nctas, ntiles);
    std::vector<float> partials((size_t)nctas * p.n_partial, NAN);
    for (int cta = 0; cta < nctas; ++cta) {
        bool first = true;
        for (long long tile = cta; tile < ntiles; tile += nctas) {
            bwd_tile<TM>(T, S.data(), z, c, W, dz, dl, x_rec, dx, dc, partials.data() + (size_t)cta * p.n_partial, first, B, tile * TM);
            first = false;
        }
    }
    for (int64_t i = 0; i < p.n_params; ++i) {
        float a = 0.f;
        for (int q = 0; q < nctas; ++q) a += partials[(size_t)q * p.n_partial + p.unpack_src[i]];
        dparams[i] = a;
    }
}
}  // namespace

extern "C" {
// returns 0 on success; tm_out receives {fwd TM, bwd TM, fwd smem, bwd smem, n_stages fwd, n_stages bwd}
int emul_run(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits, int min_split_size,
             const float* params, const float* x, const float* c, long long B, int rev, int nctas,
             float* z, float* logdet, const float* dz, const float* dl, float* x_rec, float* dx, float* dcond,
             float* dparams, long long* info) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    info[0] = p.fwd.TM; info[1] = p.bwd.TM; info[2] = (long long)p.fwd.smem_bytes; info[3] = (long long)p.bwd.smem_bytes;
    info[4] = (long long)p.fwd.stages.size(); info[5] = (long long)p.bwd.stages.size(); info[6] = p.n_params; info[7] = p.n_packed;
    std::vector<float> W = pack(p, params);
    switch (p.fwd.TM) {
        case 128: fwd_all<128>(p, x, c, W.data(), z, logdet, B, rev); break;
        case 64: fwd_all<64>(p, x, c, W.data(), z, logdet, B, rev); break;
        case 32: fwd_all<32>(p, x, c, W.data(), z, logdet, B, rev); break;
        case 16: fwd_all<16>(p, x, c, W.data(), z, logdet, B, rev); break;
        case 8: fwd_all<8>(p, x, c, W.data(), z, logdet, B, rev); break;
        default: return 100;
    }
    if (dz) {  // backward of the forward direction, fed with the z just computed
        switch (p.bwd.TM) {
            case 128: bwd_all<128>(p, nctas, z, c, W.data(), dz, dl, x_rec, dx, dcond, dparams, B); break;
            case 64: bwd_all<64>(p, nctas, z, c, W.data(), dz, dl, x_rec, dx, dcond, dparams, B); break;
            case 32: bwd_all<32>(p, nctas, z, c, W.data(), dz, dl, x_rec, dx, dcond, dparams, B); break;
            case 16: bwd_all<16>(p, nctas, z, c, W.data(), dz, dl, x_rec, dx, dcond, dparams, B); break;
            case 8: bwd_all<8>(p, nctas, z, c, W.data(), dz, dl, x_rec, dx, dcond, dparams, B); break;
            default: return 101;
        }
    }
    return 0;
}
}
