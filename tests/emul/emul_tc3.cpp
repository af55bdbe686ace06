// CPU interpreter of the tcgen05 TRAINING kernel's static programs (plan_tc3.h).  TEST INFRASTRUCTURE ONLY.
//
// TMEM is a [128][512] float array, the images / state are float arrays with the kernel's own index functions, a
// tcgen05.mma record is a plain matrix product over the packed canonical weight slab (TS) or over two images (SS).
// The two roles (MMA issuer, epilogue) are stepped as two program counters that honour the planner's inferred waits.
// Two schedules bracket what the hardware may do:
//   eager : the issuer runs as far ahead as its waits allow and every MMA completes the moment it is issued
//           (catches a missing "wait for the epilogue" on the issuer side);
//   lazy  : an MMA completes only when an epilogue step waits for its signal (or at the end of the tile)
//           (catches a missing "wait for the tensor pipe" on the epilogue side, incl. write-after-read hazards).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../hint_b200/csrc/plan_tc3.h"

using namespace hint;

namespace {
inline float rna_tf32(float x) { unsigned u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
inline float trunc_tf32(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
}

// info: [0] ok, [1] groups, [2] mma records, [3] epilogue steps, [4] chunks, [5] packed floats, [6] smem bytes,
//       [7] partial floats, [8] mma instructions per tile, [9] tensor cycles per tile, [10] hidden images, [11] slots
static int run_tc3(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits,
                   int min_split_size, const float* params, const float* z, const float* c, const float* dz,
                   const float* dJ, long long B, int lazy, int tf32, float* xrec, float* dx, float* dcond,
                   float* dparams, long long* info, int n_epi_limit, float* tmem_out, float* img_out, float* xs_out, float* gs_out,
                   float* os_out, int kind = T3K_BACKWARD, float* logdet_out = nullptr) {
    Plan p;
    int code = 0;
    std::string err = build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    if (!err.empty()) return code ? code : 1;
    T3Plan t;
    build_tc3_plan(p, t, kind);
    info[0] = t.ok; info[1] = (long long)t.groups.size(); info[2] = (long long)t.mmas.size(); info[3] = (long long)t.epis.size();
    info[4] = (long long)t.chunks.size(); info[5] = t.n_packed; info[6] = t.smem_bytes; info[7] = t.n_partial;
    info[8] = t.n_mma_instr; info[9] = t.tensor_cycles; info[10] = t.n_imgs_hidden; info[11] = t.n_slots;
    if (!t.ok) return 200;
    std::vector<float> W((size_t)t.n_packed);
    for (long long i = 0; i < t.n_packed; ++i) {
        if (t.pack_src[i] < 0) { W[i] = 0.f; continue; }
        const float v = params[t.pack_src[i] & ~kT3BiasLo];
        const float hi = tf32 ? rna_tf32(v) : v;
        W[i] = (t.pack_src[i] & kT3BiasLo) ? (tf32 ? rna_tf32(v - hi) : 0.f) : hi;
    }
    auto rt = [&](float v) { return tf32 ? rna_tf32(v) : v; };
    auto opnd = [&](float v) { return tf32 ? trunc_tf32(v) : v; };   // what the tensor core reads
    const float alpha = p.alpha;
    const int xp = t.xp, op = t.op, nd = p.d + p.dc;
    std::vector<float> T((size_t)128 * 512), XS((size_t)128 * xp), GS((size_t)128 * xp), OS((size_t)128 * op), DJ(128);
    std::vector<std::vector<float>> img(kT3Imgs);
    for (int i = 0; i < kT3Imgs; ++i) img[i].assign((size_t)std::max(8, t.img_rows[i]) * 128, NAN);
    std::vector<double> part((size_t)t.n_partial, 0.0);
    auto tm = [&](int lane, int col) -> float& { return T[(size_t)lane * 512 + col]; };
    auto im = [&](int i, int r, int s) -> float& { return img[i][(size_t)t3_img_off(r, s, t.img_rows[i])]; };
    const int nm_all = (int)t.mmas.size();
    int nm = nm_all, ne = (int)t.epis.size();
    std::vector<int> sig(nm);
    {
        int s = 0;
        for (int i = 0; i < nm; ++i) { sig[i] = s; if (t.mmas[i].flags & T3M_COMMIT) ++s; }
    }
    std::vector<int> chunk_of(nm, -1);
    {
        int ck = -1;
        for (int i = 0; i < nm; ++i) {
            if (t.mmas[i].flags & T3M_SS) continue;
            if (t.mmas[i].flags & T3M_NEWCHUNK) ++ck;
            chunk_of[i] = ck;
        }
        if (ck + 1 != (int)t.chunks.size()) return 300;
    }
    if (n_epi_limit > 0 && n_epi_limit < ne) {   // developer aid: same truncation rule as tc3_debug_run (tc3_launch.cu)
        ne = n_epi_limit;
        int m = 0;
        while (m < nm_all && t.mmas[m].wait_epi < ne) ++m;
        while (m < nm_all && m > 0 && !(t.mmas[m - 1].flags & T3M_SS) && !(t.mmas[m - 1].flags & T3M_ENDCHUNK)) ++m;
        nm = m;
        info[12] = t.epis[ne - 1].type; info[13] = t.epis[ne - 1].wait_mma; info[14] = nm; info[15] = t.epis[ne - 1].a;
    }

    auto exec_mma = [&](int i) {
        const T3Mma& m = t.mmas[i];
        const int N = (int)((m.idesc >> 17) & 0x3F) << 3;
        std::vector<double> acc((size_t)128 * N, 0.0);
        if (m.flags & T3M_SS) {
            const int ai = m.b_off & 0xFF, bi = (m.b_off >> 8) & 0xFF;
            for (int j = 0; j < 128; ++j) {
                const int r = m.a_col * 8 + j;
                if (r >= t.img_rows[ai]) { for (int n = 0; n < N; ++n) acc[(size_t)j * N + n] = NAN; continue; }   // over-read rows: never flushed
                for (int n = 0; n < N; ++n) {
                    if (n >= t.img_rows[bi]) { acc[(size_t)j * N + n] = NAN; continue; }   // B rows past the allocation: columns no parameter maps to
                    double a = 0;
                    for (int s = 0; s < 128; ++s) a += (double)opnd(im(ai, r, s)) * (double)opnd(im(bi, n, s));
                    acc[(size_t)j * N + n] = a;
                }
            }
        } else {
            const T3Chunk& ck = t.chunks[chunk_of[i]];
            const int Kc = m.b_sbo16 / 2;
            const float* Bm = W.data() + ck.g_off + m.b_off / 4;
            for (int lane = 0; lane < 128; ++lane)
                for (int n = 0; n < N; ++n) {
                    double a = 0;
                    for (int k = 0; k < 8 * m.nk; ++k)
                        a += (double)opnd(tm(lane, m.a_col + k)) * (double)Bm[(n >> 3) * (Kc * 8) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)];
                    acc[(size_t)lane * N + n] = a;
                }
        }
        for (int lane = 0; lane < 128; ++lane)
            for (int n = 0; n < N; ++n) {
                float& dref = tm(lane, m.d_col + n);
                dref = (m.flags & T3M_ZERO) ? (float)acc[(size_t)lane * N + n] : (float)((double)dref + acc[(size_t)lane * N + n]);
            }
    };

    for (long long row0 = 0; row0 < B; row0 += 128) {
        const int rows = (int)std::min<long long>(128, B - row0);
        for (auto& v : T) v = NAN;
        for (int s = 0; s < 128; ++s) {
            for (int j = 0; j < nd; ++j) {
                float xv = 0.f, gv = 0.f;
                if (s < rows) {
                    if (j < p.d) { xv = z[(row0 + s) * p.d + j]; gv = dz ? dz[(row0 + s) * p.d + j] : 0.f; }
                    else xv = c[(row0 + s) * p.dc + (j - p.d)];
                }
                XS[(size_t)s * xp + j] = xv; GS[(size_t)s * xp + j] = gv;
            }
            DJ[s] = (s < rows && dJ) ? dJ[row0 + s] : 0.f;
        }
        std::vector<float> JA(128, 0.f);   // transport kernels: log-det accumulators
        int pm = 0, pe = 0, done_sig = -1, issued_upto = 0;   // records [issued_upto, pm) are issued but not executed (lazy)
        auto run_queued = [&](int upto_sig) {
            while (issued_upto < pm && sig[issued_upto] <= upto_sig) exec_mma(issued_upto++);
        };
        auto epi_step = [&](const T3Epi& e) {
            const int16_t* tb = t.tab16.data();
            switch (e.type) {
                case T3E_IN:
                    for (int s = 0; s < 128; ++s)
                        for (int col = 0; col < e.b; ++col) {
                            const int code = tb[e.a + col];
                            const float v = rt(code >= 0 ? XS[(size_t)s * xp + code] : (code == -2 ? 1.f : 0.f));
                            tm(s, e.c + col) = v; if (!(e.flags & T3I_NOIMG)) im(3, col, s) = v;
                        }
                    break;
                case T3E_HID:
                    for (int s = 0; s < 128; ++s) {
                        for (int col = 0; col < e.b; ++col) {
                            const float v = rt(std::max(tm(s, e.a + col), 0.f));
                            tm(s, e.a + col) = v;
                            if (e.flags & T3H_IMG) im(e.c, col, s) = v;
                        }
                        if (e.flags & T3H_ONES) for (int q = 0; q < 8; ++q) tm(s, e.a + e.b + q) = q < 2 ? 1.f : 0.f;
                        if (e.flags & T3H_IMG_ONES) for (int q = 0; q < 8; ++q) im(e.c, e.b + q, s) = q < 2 ? 1.f : 0.f;
                    }
                    break;
                case T3E_OUTS:
                    for (int s = 0; s < 128; ++s)
                        for (int col = 0; col < e.b; ++col) OS[(size_t)s * op + col] = tm(s, e.a + col);
                    break;
                case T3E_CPL:
                    for (int col = 0; col < e.b; ++col) {
                        double bsum = 0;
                        const int xc = tb[e.c + col];
                        for (int s = 0; s < 128; ++s) {
                            const float sv = OS[(size_t)s * op + col], tv = tm(s, e.a + col);
                            const float zl = XS[(size_t)s * xp + xc], dzl = GS[(size_t)s * xp + xc];
                            const float la = alpha * atanf(sv), ee = expf(la);
                            const float xl = (zl - tv) / ee;
                            XS[(size_t)s * xp + xc] = xl;
                            GS[(size_t)s * xp + xc] = dzl * ee;
                            const float ds = (dzl * xl * ee + DJ[s]) * alpha / (1.f + sv * sv);
                            OS[(size_t)s * op + col] = ds;
                            tm(s, e.d + col) = rt(dzl); im(4, col, s) = rt(dzl);
                            bsum += dzl;
                            if ((s & 31) == 31) { part[(size_t)e.off + 32 * (s >> 5) + col] += bsum; bsum = 0; }
                        }
                    }
                    for (int s = 0; s < 128; ++s) for (int col = e.b; col < e.e; ++col) tm(s, e.d + col) = 0.f;
                    break;
                case T3E_CPLF:
                    for (int col = 0; col < e.b; ++col) {
                        const int xc = tb[e.c + col];
                        for (int s = 0; s < 128; ++s) {
                            const float la = alpha * atanf(OS[(size_t)s * op + col]), tv = tm(s, e.a + col);
                            float& xr = XS[(size_t)s * xp + xc];
                            if (!e.flags) { xr = expf(la) * xr + tv; JA[s] += la; }
                            else { xr = (xr - tv) / expf(la); JA[s] -= la; }
                        }
                    }
                    break;
                case T3E_DS:
                    for (int col = 0; col < e.b; ++col) {
                        double bsum = 0;
                        for (int s = 0; s < 128; ++s) {
                            const float ds = OS[(size_t)s * op + col];
                            tm(s, e.a + col) = rt(ds); im(4, col, s) = rt(ds);
                            bsum += ds;
                            if ((s & 31) == 31) { part[(size_t)e.off + 32 * (s >> 5) + col] += bsum; bsum = 0; }
                        }
                    }
                    for (int s = 0; s < 128; ++s) for (int col = e.b; col < e.e; ++col) tm(s, e.a + col) = 0.f;
                    break;
                case T3E_DHID:
                    for (int s = 0; s < 128; ++s)
                        for (int col = 0; col < e.b; ++col) {
                            const float hv = (e.flags & T3D_MASK_TMEM) ? tm(s, e.e + col) : im(e.e, col, s);
                            const float v = hv > 0.f ? rt(tm(s, e.a + col)) : 0.f;
                            tm(s, e.a + col) = v; im(e.c, col, s) = v;
                        }
                    break;
                case T3E_DA:
                    for (int s = 0; s < 128; ++s)
                        for (int col = 0; col < e.b; ++col) GS[(size_t)s * xp + tb[e.c + col]] += tm(s, e.a + col);
                    break;
                case T3E_FLUSH: {
                    const int nn = tb[e.f];
                    for (int lane = 0; lane < e.e; ++lane) {
                        const int row = e.h * 128 + lane;
                        int c0 = 0, c1 = 0; bool found = false;
                        for (int q = 0; q < nn; ++q) {
                            const int16_t* nt = tb + e.f + 1 + 6 * q;
                            if (row >= nt[0] && row < nt[0] + nt[1]) {
                                found = true;
                                if (e.g == T3F_W2) { c0 = nt[0]; c1 = nt[0] + nt[1]; }
                                else if (e.g == T3F_W1) { c0 = nt[2]; c1 = nt[2] + nt[3]; }
                                else { c0 = nt[4]; c1 = nt[4] + nt[5]; }
                            }
                        }
                        if (!found) continue;
                        const int x0 = e.c, x1 = e.g == T3F_W2 ? e.c + 1 : e.g == T3F_W1 ? e.c + p.dc + 1 : e.c;
                        for (int col = 0; col < e.b; ++col)
                            if ((col >= c0 && col < c1) || (e.c >= 0 && col >= x0 && col < x1))
                                part[(size_t)e.off + (size_t)col * 128 + lane] += tm(lane, e.a + col);
                    }
                    break;
                }
                default: break;
            }
        };
        long long guard = 0;
        while (pm < nm || pe < ne) {
            if (++guard > 4LL * (nm + ne) + 16) return 400;   // deadlock in the inferred waits
            // issuer: as far as its waits allow
            while (pm < nm && t.mmas[pm].wait_epi < pe) {
                ++pm;
                if (!lazy) { exec_mma(issued_upto++); }
            }
            if (!lazy) done_sig = pm > 0 ? sig[pm - 1] - ((t.mmas[pm - 1].flags & T3M_COMMIT) ? 0 : 1) : -1;
            // epilogue: one step
            if (pe < ne) {
                const T3Epi& e = t.epis[pe];
                // the signal must have been issued
                const int issued_sig = pm > 0 ? sig[pm - 1] - ((t.mmas[pm - 1].flags & T3M_COMMIT) ? 0 : 1) : -1;
                if (e.wait_mma <= issued_sig) {
                    if (lazy) run_queued(e.wait_mma);
                    epi_step(e);
                    ++pe;
                }
            }
        }
        if (lazy) run_queued(1 << 30);
        (void)done_sig;
        for (int s = 0; s < rows; ++s) {
            if (kind != T3K_BACKWARD) {
                for (int j = 0; j < p.d; ++j) dx[(row0 + s) * p.d + j] = XS[(size_t)s * xp + j];
                logdet_out[row0 + s] = JA[s];
                continue;
            }
            for (int j = 0; j < p.d; ++j) {
                dx[(row0 + s) * p.d + j] = GS[(size_t)s * xp + j];
                if (xrec) xrec[(row0 + s) * p.d + j] = XS[(size_t)s * xp + j];
            }
            for (int j = 0; j < p.dc; ++j) dcond[(row0 + s) * p.dc + j] = GS[(size_t)s * xp + p.d + j];
        }
    }
    if (kind != T3K_BACKWARD) return 0;
    for (long long i = 0; i < p.n_params; ++i) {
        double v = part[(size_t)t.unpack_src[i]];
        if (t.unpack_q4[(size_t)i]) for (int q = 1; q < 4; ++q) v += part[(size_t)t.unpack_src[i] + 32 * q];
        dparams[i] = (float)v;
    }
    if (tmem_out) {
        memcpy(tmem_out, T.data(), T.size() * 4);
        size_t o = 0;
        for (int i = 0; i < kT3Imgs; ++i) { memcpy(img_out + o, img[i].data(), (size_t)t.img_rows[i] * 128 * 4); o += (size_t)t.img_rows[i] * 128; }
        memcpy(xs_out, XS.data(), XS.size() * 4); memcpy(gs_out, GS.data(), GS.size() * 4); memcpy(os_out, OS.data(), OS.size() * 4);
    }
    return 0;
}

extern "C" int emul_tc3_backward(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits,
                                 int min_split_size, const float* params, const float* z, const float* c, const float* dz,
                                 const float* dJ, long long B, int lazy, int tf32, float* xrec, float* dx, float* dcond,
                                 float* dparams, long long* info) {
    return run_tc3(d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, params, z, c, dz, dJ, B, lazy, tf32, xrec, dx, dcond,
                   dparams, info, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// forward (rev = 0) / inverse (rev = 1) transport through the T3K_FORWARD / T3K_INVERSE programs
extern "C" int emul_tc3_transport(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits, int min_split_size,
                                  const float* params, const float* x, const float* c, long long B, int rev, int lazy, int tf32,
                                  float* z, float* logdet, long long* info) {
    return run_tc3(d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, params, x, c, nullptr, nullptr, B, lazy, tf32, nullptr, z,
                   nullptr, nullptr, info, 0, nullptr, nullptr, nullptr, nullptr, nullptr, rev ? T3K_INVERSE : T3K_FORWARD, logdet);
}

// developer aid: ONE tile (B <= 128), stop after n_epi_limit epilogue steps, dump TMEM [128][512], the 5 images (rows*128 floats
// each, kernel layout), and the state arrays [128][xp], [128][xp], [128][op]
extern "C" int emul_tc3_debug(int d, int dc, const int* c_internal, int n_internal, double clamp, int max_splits,
                              int min_split_size, const float* params, const float* z, const float* c, const float* dz,
                              const float* dJ, long long B, int tf32, int n_epi_limit, float* tmem_out, float* img_out, float* xs_out,
                              float* gs_out, float* os_out, long long* info) {
    std::vector<float> xrec((size_t)B * d), dx((size_t)B * d), dcond((size_t)B * (dc ? dc : 1));
    Plan p; int code = 0;
    build_plan(p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, 0, &code);
    std::vector<float> dparams((size_t)p.n_params);
    return run_tc3(d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, params, z, c, dz, dJ, B, 0, tf32, xrec.data(), dx.data(),
                   dcond.data(), dparams.data(), info, n_epi_limit, tmem_out, img_out, xs_out, gs_out, os_out);
}
