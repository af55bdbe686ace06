// CPU interpreter of the tcgen05 (TF32) kernel's static program (plan_tc.h).  TEST INFRASTRUCTURE ONLY.
// TMEM is a [128][512] float array, an MMA op is a plain matrix product reading the packed canonical
// weight image exactly as the tensor core would; epilogues follow the kernel.  Validates the schedule
// generator, the packed layouts and the column bookkeeping without a GPU.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../hint_b200/csrc/plan_tc.h"

using namespace hint;

namespace {
inline int canon_off(int r, int k, int kpad) { return (r >> 3) * (kpad * 8) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }
inline float trunc_tf32(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
inline float rna_tf32(float x) { unsigned u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
}

extern "C" int emul_tc_run(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits,
                           int min_split_size, const float* params, const float* x, const float* c, long long B, int rev,
                           int tf32, float* z, float* logdet, long long* info) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    TcSchedule t;
    build_tc_schedule(p, t);
    info[0] = t.ok; info[1] = t.xw; info[2] = t.xr; info[3] = (long long)t.stages.size(); info[4] = (long long)t.ops.size();
    info[5] = (long long)t.chunks.size(); info[6] = t.n_packed; info[7] = (long long)t.smem_bytes; info[8] = t.slot_bytes; info[9] = t.n_slots;
    if (!t.ok) return 200;
    std::vector<float> W((size_t)t.n_packed);
    for (long long i = 0; i < t.n_packed; ++i) {
        float v = t.pack_src[i] < 0 ? 0.f : params[t.pack_src[i]];
        W[i] = (tf32 && i < t.n_weight_floats) ? rna_tf32(v) : v;
    }
    const float alpha = p.alpha;
    std::vector<float> T((size_t)128 * 512);
    auto at = [&](int m, int col) -> float& { return T[(size_t)m * 512 + col]; };
    for (long long row0 = 0; row0 < B; row0 += 128) {
        for (auto& v : T) v = NAN;
        const int rows = (int)std::min<long long>(128, B - row0);
        std::vector<float> J(128, 0.f);
        for (int m = 0; m < 128; ++m) {
            for (int pc = 0; pc < t.xr; ++pc) {
                float v = 0.f;
                if (m < rows) {
                    if (pc < t.xw) { if (t.xlog[pc] >= 0) v = x[(row0 + m) * d + t.xlog[pc]]; }
                    else if (pc - t.xc < dc) v = c[(row0 + m) * dc + (pc - t.xc)];
                }
                at(m, pc) = v;
            }
        }
        const int S = (int)t.stages.size();
        for (int si = 0; si < S; ++si) {
            const TcStage& st = t.stages[rev ? si : S - 1 - si];
            bool epi_done[3] = {false, false, false};
            auto hidden = [&](int j) {
                if (epi_done[j]) return;
                epi_done[j] = true;
                const TcHidden& h = st.hid[j];
                for (int m = 0; m < 128; ++m)
                    for (int q = 0; q < h.ncols; ++q) {
                        float v = at(m, h.col0 + q) + W[h.bias_off + q];
                        v = v > 0.f ? v : 0.f;
                        at(m, h.col0 + q) = tf32 ? rna_tf32(v) : v;
                    }
            };
            int chunk = st.chunk_begin - 1;
            for (int oi = st.op_begin; oi < st.op_end; ++oi) {
                const TcOp& op = t.ops[oi];
                if (op.flags & TC_FIRST_IN_CHUNK) ++chunk;
                if (chunk < st.chunk_begin || chunk >= st.chunk_end) return 201;
                if (op.wait_epi >= 0) hidden(op.wait_epi);
                const float* img = W.data() + t.chunks[chunk].g_off + op.b_off;
                if ((op.b_off * 4 + op.n_rows * op.nk * 32) > t.chunks[chunk].bytes) return 202;
                const int kpad = 8 * op.nk;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < op.n_rows; ++n) {
                        float acc = (op.flags & TC_ACCUM) ? at(m, op.d_col + n) : 0.f;
                        for (int k = 0; k < kpad; ++k) {
                            float a = at(m, op.a_col + k);
                            if (tf32) a = trunc_tf32(a);
                            acc += a * img[canon_off(n, k, kpad)];
                        }
                        at(m, op.d_col + n) = acc;
                    }
            }
            for (int fi = st.fin_begin; fi < st.fin_end; ++fi) {
                const TcFinal& f = t.fins[fi];
                for (int m = 0; m < 128; ++m)
                    for (int q = 0; q < 4; ++q) {
                        const float s = at(m, f.s_col + q) + W[f.bs_off + q];
                        const float tt = at(m, f.t_col + q) + W[f.bt_off + q];
                        const float la = alpha * atanf(s);
                        float& xv = at(m, f.x_col + q);
                        if (!rev) { xv = expf(la) * xv + tt; J[m] += la; }
                        else { xv = (xv - tt) / expf(la); J[m] -= la; }
                    }
            }
        }
        for (int m = 0; m < rows; ++m) {
            for (int cc = 0; cc < d; ++cc) z[(row0 + m) * d + cc] = at(m, t.xphys[cc]);
            logdet[row0 + m] = J[m];
        }
    }
    return 0;
}

// ---- v2 encoding (T2Prog, tc2_kernels.cuh): same machine, ops fetched per segment / issuer --------------------
extern "C" int emul_tc2_run(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits,
                            int min_split_size, const float* params, const float* x, const float* c, long long B, int rev,
                            int tf32, float* z, float* logdet, long long* info) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    TcSchedule t;
    build_tc_schedule(p, t);
    static T2Host H;   // ~28 KB
    build_tc2_program(p, t, H);
    info[0] = H.ok; info[1] = (long long)H.smem_bytes; info[2] = H.ok ? H.prog.n_slots : 0; info[3] = (long long)sizeof(T2Prog);
    if (!H.ok) return 200;
    const T2Prog& P = H.prog;
    std::vector<float> W((size_t)t.n_packed);
    for (long long i = 0; i < t.n_packed; ++i) {
        float v = t.pack_src[i] < 0 ? 0.f : params[t.pack_src[i]];
        W[i] = (tf32 && i < t.n_weight_floats) ? rna_tf32(v) : v;
    }
    const float* bias = W.data() + P.bias_base;
    std::vector<float> T((size_t)128 * 512);
    auto at = [&](int m, int col) -> float& { return T[(size_t)m * 512 + col]; };
    for (long long row0 = 0; row0 < B; row0 += 128) {
        for (auto& v : T) v = NAN;
        const int rows = (int)std::min<long long>(128, B - row0);
        std::vector<float> J(128, 0.f);
        for (int m = 0; m < 128; ++m)
            for (int pc = 0; pc < P.xr; ++pc) {
                float v = 0.f;
                if (m < rows) {
                    if (pc < P.xw) { if (P.xlog[pc] >= 0) v = x[(row0 + m) * d + P.xlog[pc]]; }
                    else if (pc - P.xc < dc) v = c[(row0 + m) * dc + (pc - P.xc)];
                }
                at(m, pc) = v;
            }
        for (int si = 0; si < P.nstages; ++si) {
            const T2Stage& S = P.stages[rev ? si : P.nstages - 1 - si];
            bool epi_done[3] = {false, false, false};
            auto hidden = [&](int j) {
                if (epi_done[j]) return;
                epi_done[j] = true;
                const T2Hidden& h = S.hid[j];
                for (int m = 0; m < 128; ++m)
                    for (int q = 0; q < h.ncols; ++q) {
                        float v = at(m, h.col0 + q) + bias[h.bias_off + q];
                        v = v > 0.f ? v : 0.f;
                        at(m, h.col0 + q) = tf32 ? rna_tf32(v) : v;
                    }
            };
            int chunk = S.chunk_begin;
            int jobs_seen = 0;
            for (int g = S.seg_begin; g < S.seg_end; ++g) {
                const T2Seg& G = P.segs[g];
                if (G.flags & T2_FIRST_IN_JOB) {
                    if (G.job != jobs_seen) return 210;   // every job exactly once, in order
                    const int dep = G.job == TC_J1 ? -1 : (G.job == TC_J2S || G.job == TC_J2T) ? 0 : (G.job == TC_J3S ? 1 : 2);
                    if (dep >= 0) hidden(dep);
                }
                for (int oi = G.op_ofs[0]; oi < G.op_ofs[kTcIssuers]; ++oi) {
                    if (chunk >= S.chunk_end) return 211;
                    const T2Op& op = P.ops[oi];
                    const int d_col = op.da & 0xFFFF, a_col = op.da >> 16, nk = op.sbo_nk >> 16, kpad = 8 * nk;
                    const int n_rows = ((op.idesc >> 17) & 0x3F) << 3;
                    if ((int)(op.sbo_nk & 0xFFFF) * 16 != nk * 256) return 212;
                    if (op.b16 * 16 + (unsigned)(n_rows * kpad * 4) > P.chunks[chunk].bytes) return 213;
                    const float* img = W.data() + (size_t)P.chunks[chunk].g_off16 * 4 + (size_t)op.b16 * 4;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < n_rows; ++n) {
                            float acc = (op.idesc & 1u) ? at(m, d_col + n) : 0.f;
                            for (int k = 0; k < kpad; ++k) {
                                float a = at(m, a_col + k);
                                if (tf32) a = trunc_tf32(a);
                                acc += a * img[canon_off(n, k, kpad)];
                            }
                            at(m, d_col + n) = acc;
                        }
                }
                if (G.flags & T2_LAST_IN_CHUNK) ++chunk;
                if (G.flags & T2_LAST_IN_JOB) ++jobs_seen;
            }
            if (jobs_seen != TC_NJOBS || chunk != S.chunk_end) return 214;
            for (int fi = S.fin_begin; fi < S.fin_end; ++fi) {
                const T2Fin& f = P.fins[fi];
                for (int m = 0; m < 128; ++m)
                    for (int q = 0; q < 4; ++q) {
                        const float s = at(m, f.s_col + q) + bias[f.bs_off + q];
                        const float tt = at(m, f.t_col + q) + bias[f.bt_off + q];
                        const float la = P.alpha * atanf(s);
                        float& xv = at(m, f.x_col + q);
                        if (!rev) { xv = expf(la) * xv + tt; J[m] += la; }
                        else { xv = (xv - tt) / expf(la); J[m] -= la; }
                    }
            }
        }
        for (int m = 0; m < rows; ++m) {
            for (int cc = 0; cc < d; ++cc) z[(row0 + m) * d + cc] = at(m, t.xphys[cc]);
            logdet[row0 + m] = J[m];
        }
    }
    return 0;
}
