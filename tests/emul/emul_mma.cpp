// Host emulation of the warp-MMA kernels (mma_kernels.cuh).  TEST INFRASTRUCTURE ONLY: built and loaded by
// tests/test_emul_mma.py on CPU-only machines; never linked into libhint_b200.so, never used by the product.
//
// A CTA is emulated as kMmaThreads cooperative fibers (ucontext) scheduled round-robin on one host thread.  The CTA
// barrier and the warp-collective MMA are rendezvous points: a fiber that arrives early yields until the others have
// arrived.  mma_tf32 exchanges the lanes' fragments through a per-warp buffer, truncates the operands to 10 mantissa
// bits (what the tensor core does with tf32 inputs) and evaluates the PTX m16n8k8 fragment layout, so the planner's
// packing, the k-slot permutation, the task tables and every index of the kernels are exercised without a GPU.
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

using std::min;
#include "../../hint_b200/csrc/plan.h"
#include "../../hint_b200/csrc/plan_mma.h"
#include "../../hint_b200/csrc/mma_kernels.cuh"

namespace hint { namespace emu {

struct WarpX {
    int count = 0;
    unsigned gen = 0;
    const float* rowp[32];
    float shv[32];
    uint32_t a[32][4];
    uint32_t b[32][2];
};

struct Cta {
    int n = 0, cur = 0;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stacks;
    std::vector<char> done;
    ucontext_t main;
    int bar_count = 0;
    unsigned bar_gen = 0;
    WarpX warps[32];
    void (*body)(int tid, void* arg) = nullptr;
    void* arg = nullptr;
    long long switches = 0;
};

static Cta* g_cta = nullptr;

static void yield_fiber() {
    Cta* c = g_cta;
    ++c->switches;
    swapcontext(&c->ctx[c->cur], &c->main);
}

void cta_sync() {
    Cta* c = g_cta;
    const unsigned gen = c->bar_gen;
    if (++c->bar_count == c->n) { c->bar_count = 0; ++c->bar_gen; }
    else while (c->bar_gen == gen) yield_fiber();
}

static void warp_rendezvous(WarpX& w) {
    const unsigned gen = w.gen;
    if (++w.count == 32) { w.count = 0; ++w.gen; }
    else while (w.gen == gen) yield_fiber();
}

void warp_sync() { warp_rendezvous(g_cta->warps[g_cta->cur >> 5]); }

// ldmatrix.m8n8.x4 on 32-bit elements: lane l supplies row l%8 of matrix l/8; r[m] = element (row lane/4, column lane%4)
void ldsm4(const float* rowp, uint32_t (&r)[4]) {
    Cta* cta = g_cta;
    const int lane = cta->cur & 31;
    WarpX& w = cta->warps[cta->cur >> 5];
    w.rowp[lane] = rowp;
    warp_rendezvous(w);
    for (int m = 0; m < 4; ++m) std::memcpy(&r[m], w.rowp[8 * m + (lane >> 2)] + (lane & 3), 4);
    warp_rendezvous(w);
}

float shfl_xor(float v, int mask) {
    Cta* cta = g_cta;
    const int lane = cta->cur & 31;
    WarpX& w = cta->warps[cta->cur >> 5];
    w.shv[lane] = v;
    warp_rendezvous(w);
    const float r = w.shv[lane ^ mask];
    warp_rendezvous(w);
    return r;
}

static inline float tf32(uint32_t u) {
    u &= 0xFFFFE000u;
    float v;
    std::memcpy(&v, &u, 4);
    return v;
}

// PTX mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 fragment layout (g = lane/4, t = lane%4):
//   A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (k=t, n=g) b1 (k=t+4, n=g);  C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    Cta* cta = g_cta;
    const int lane = cta->cur & 31;
    WarpX& w = cta->warps[cta->cur >> 5];
    for (int e = 0; e < 4; ++e) w.a[lane][e] = a[e];
    w.b[lane][0] = b0; w.b[lane][1] = b1;
    warp_rendezvous(w);
    const int g = lane >> 2, t = lane & 3;
    for (int e = 0; e < 4; ++e) {
        const int row = g + 8 * (e >> 1), col = 2 * t + (e & 1);
        float acc = c[e];
        for (int k = 0; k < 8; ++k) {
            const float av = tf32(w.a[(row & 7) * 4 + (k & 3)][(row >> 3) + 2 * (k >> 2)]);
            const float bv = tf32(w.b[col * 4 + (k & 3)][k >> 2]);
            acc = fmaf(av, bv, acc);
        }
        c[e] = acc;
    }
    warp_rendezvous(w);
}

static void trampoline() {
    Cta* c = g_cta;
    const int tid = c->cur;
    c->body(tid, c->arg);
    c->done[tid] = 1;
    swapcontext(&c->ctx[tid], &c->main);
}

void run_cta(int nthreads, void (*body)(int, void*), void* arg) {
    Cta cta;
    cta.n = nthreads;
    cta.ctx.resize(nthreads);
    cta.stacks.assign(nthreads, std::vector<char>(96 * 1024));
    cta.done.assign(nthreads, 0);
    cta.body = body;
    cta.arg = arg;
    g_cta = &cta;
    for (int i = 0; i < nthreads; ++i) {
        getcontext(&cta.ctx[i]);
        cta.ctx[i].uc_stack.ss_sp = cta.stacks[i].data();
        cta.ctx[i].uc_stack.ss_size = cta.stacks[i].size();
        cta.ctx[i].uc_link = &cta.main;
        makecontext(&cta.ctx[i], trampoline, 0);
    }
    for (;;) {
        bool any = false;
        for (int i = 0; i < nthreads; ++i) {
            if (cta.done[i]) continue;
            any = true;
            cta.cur = i;
            swapcontext(&cta.main, &cta.ctx[i]);
        }
        if (!any) break;
    }
    g_cta = nullptr;
}

} }  // namespace hint::emu

using namespace hint;

namespace {

MmaTables tables(const Plan& p, const MSchedule& s, int prog) {
    MmaTables t;
    t.prog = s.prog.data(); t.eps = s.eps.data();
    t.in_param = s.fits_param[prog] ? 1 : 0;   // exercise the same record source the GPU launch would use
    for (int w = 0; w < kMmaWarps; ++w) t.begin[w] = s.prog_begin[prog][w] - (t.in_param ? s.prog_begin[prog][0] : 0);
    t.d = p.d; t.dc = p.dc;
    t.col_x = s.col_x; t.col_d = s.col_d; t.col_one = s.col_one; t.col_zero = s.col_zero;
    t.raw_off = s.raw_off; t.alpha = p.alpha; t.dbg = nullptr; t.wcopies = 1; t.wstride = 0;
    return t;
}

void fill_param_prog(const MSchedule& s, int prog, MmaParamProg& P) {
    std::memset(&P, 0, sizeof(P));
    if (!s.fits_param[prog]) return;
    const int b = s.prog_begin[prog][0], n = s.prog_end[prog] - b;
    std::memcpy(P.ops, s.prog.data() + b, (size_t)n * sizeof(WOp));
    for (size_t i = 0; i < s.eps.size(); ++i) {
        P.eps[i][0] = (unsigned short)s.eps[i].x_col; P.eps[i][1] = (unsigned short)s.eps[i].s_col;
        P.eps[i][2] = (unsigned short)s.eps[i].t_col; P.eps[i][3] = 0;
    }
}

struct FwdArgs {
    const MmaParamProg* P;
    MmaTables T; float* S; const float *x, *c, *W, *Wlo; float *z, *logdet; long long B; int rev, bid, nblocks, TM; bool x3;
};
template <int TM, bool X3>
void fwd_body(int tid, void* a) {
    FwdArgs& A = *(FwdArgs*)a;
    m_fwd_body<TM, X3>(A.T, *A.P, A.S, A.x, A.c, A.W, A.Wlo, A.z, A.logdet, A.B, A.rev, tid, A.bid, A.nblocks);
}
struct BwdArgs {
    const MmaParamProg* P;
    MmaTables T; float* S; const float *z, *c, *W, *Wlo, *dz, *dl; float *x_rec, *dx, *dc, *partials; long long n_partial, B;
    int bid, nblocks;
};
template <int TM, bool X3>
void bwd_body(int tid, void* a) {
    BwdArgs& A = *(BwdArgs*)a;
    m_bwd_body<TM, X3>(A.T, *A.P, A.S, A.z, A.c, A.W, A.Wlo, A.dz, A.dl, A.x_rec, A.dx, A.dc, A.partials, A.n_partial, A.B, tid, A.bid, A.nblocks);
}

typedef void (*BodyFn)(int, void*);
BodyFn pick_fwd(int TM, bool x3) {
    switch (TM) {
        case 64: return x3 ? fwd_body<64, true> : fwd_body<64, false>;
        case 32: return x3 ? fwd_body<32, true> : fwd_body<32, false>;
    }
    return nullptr;
}
BodyFn pick_bwd(int TM, bool x3) {
    switch (TM) {
        case 64: return x3 ? bwd_body<64, true> : bwd_body<64, false>;
        case 32: return x3 ? bwd_body<32, true> : bwd_body<32, false>;
    }
    return nullptr;
}

}  // namespace

extern "C" {
// info: {fwd TM, bwd TM, fwd smem, bwd smem, fwd stages, bwd stages, fwd ctas/SM, bwd ctas/SM, n_packed, n_partial,
//        fwd mtasks, bwd mtasks, bwd dtasks, fiber switches}
int emul_mma_run(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits, int min_split_size,
                 const float* params, const float* x, const float* c, long long B, int rev, int nctas, int x3,
                 float* z, float* logdet, const float* dz, const float* dl, float* x_rec, float* dx, float* dcond,
                 float* dparams, long long* info) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    MmaPlan m;
    build_mma_plan(p, m);
    if (!m.ok) return 50;
    info[0] = m.fwd.TM; info[1] = m.bwd.TM; info[2] = (long long)m.fwd.smem_bytes; info[3] = (long long)m.bwd.smem_bytes;
    info[4] = (long long)m.fwd.stages.size(); info[5] = (long long)m.bwd.stages.size();
    info[6] = m.fwd.ctas_per_sm; info[7] = m.bwd.ctas_per_sm; info[8] = m.n_packed; info[9] = m.n_partial;
    info[13] = m.fwd.fits_param[0] + 2 * m.bwd.fits_param[0]; info[10] = (long long)m.fwd.mtasks.size(); info[11] = (long long)m.bwd.mtasks.size(); info[12] = (long long)m.bwd.dtasks.size();
    std::vector<float> W((size_t)m.n_packed + 4), Wlo((size_t)m.n_packed + 4);
    for (int64_t i = 0; i < m.n_packed; ++i) m_pack_elem(m.pack_src[(size_t)i], params, W[(size_t)i], Wlo[(size_t)i]);
    long long switches = 0;
    {
        const MSchedule& s = m.fwd;
        BodyFn fn = pick_fwd(s.TM, x3 != 0);
        if (!fn) return 100;
        static MmaParamProg PP;
        fill_param_prog(s, rev ? PROG_INV : PROG_FWD, PP);
        const long long ntiles = (B + s.TM - 1) / s.TM;
        const int nb = (int)std::max<long long>(1, std::min<long long>(nctas, ntiles));
        for (int bid = 0; bid < nb; ++bid) {
            std::vector<float> S(s.smem_bytes / 4 + 16, NAN);
            FwdArgs A{&PP, tables(p, s, rev ? PROG_INV : PROG_FWD), S.data(), x, c, W.data(), Wlo.data(), z, logdet, B, rev, bid, nb, s.TM, x3 != 0};
            emu::run_cta(kMmaThreads, fn, &A);
        }
    }
    if (dz) {
        const MSchedule& s = m.bwd;
        BodyFn fn = pick_bwd(s.TM, x3 != 0);
        if (!fn) return 101;
        static MmaParamProg PP;
        fill_param_prog(s, PROG_BWD, PP);
        const long long ntiles = (B + s.TM - 1) / s.TM;
        const int nb = (int)std::max<long long>(1, std::min<long long>(nctas, ntiles));
        std::vector<float> partials((size_t)nb * m.n_partial, NAN);
        for (int bid = 0; bid < nb; ++bid) {
            std::vector<float> S(s.smem_bytes / 4 + 16, NAN);
            BwdArgs A{&PP, tables(p, s, PROG_BWD), S.data(), z, c, W.data(), Wlo.data(), dz, dl, x_rec, dx, dcond, partials.data(), m.n_partial, B, bid, nb};
            emu::run_cta(kMmaThreads, fn, &A);
        }
        for (int64_t i = 0; i < p.n_params; ++i) {
            float a = 0.f;
            for (int q = 0; q < nb; ++q) a += partials[(size_t)q * m.n_partial + m.unpack_src[(size_t)i]];
            dparams[i] = a;
        }
    }
    return 0;
}
}
