// The baseline couplings the 2-lane conditional HINT configs put either side of the HINT block (SURVEY.md 8f-4;
// configs/lens_shape/conditional_hint_8_full.py:78-89): FrEIA's `AffineCoupling` / `ExternalAffineCoupling` with `F_fully_connected`
// subnets (fc1-ReLU-fc2-ReLU-fc2b-ReLU-fc3).  FrEIA's source is not part of the reference: published definition, parity-unpinned;
// the checker is the plain-PyTorch statement of the same definition in FrEIA/modules/coupling.py.
//
//   y = e(s(u)) * v + t(u),  logdet = sum_c log e(s_c),  e(s) = exp(clamp * 0.636 * atan(s))      (rev: y = (v - t(u)) / e(s(u)))
//
// u [B, du] is the subnets' input (x1 | condition, or the condition alone), v [B, dv] the transformed part.  Everything is tiny per
// sample (du <= 4 ... 50, H = 17 ... 224) and was launch-bound in PyTorch (about 75 kernels per coupling and step); here:
//   mc_forward_kernel   ONE launch: a warp carries 16 samples through both four-layer subnets (activations ping-pong in its private
//                       shared-memory rows, weights straight from L1/L2 in PyTorch's [out][in] layout = the col-major B operand)
//                       and applies the coupling in the last layer's epilogue.
//   mc_backward_kernel  ONE launch: recomputes the activations, back-propagates both subnets (masked dgrad chain), writes dL/du,
//                       dL/dv and leaves the per-layer (input, output-gradient) pairs in the workspace;
//   mc_wgrad_kernel     ONE launch for all 8 layers: dW = D^T X over the batch, 8 warps split the samples, fixed-order reduction
//                       (deterministic, no atomics), bias gradient = column sums from the same fragments.
// All products are error-compensated 3 x TF32 `mma.m16n8k8` (hi/lo split of both operands, fp32 accumulation): fp32-grade, like
// the fp32 GEMMs they replace.
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <tuple>

#include "launch_count.h"
#include "mlp_coupling.h"

namespace hint {

namespace {

constexpr int kMcThreads = 512;
constexpr int kMcMaxH = 256, kMcMaxIO = 128;
constexpr int kMcWgWarps = 8;

struct McNet { const float* W[4]; const float* b[4]; };   // fc1 [H][du], fc2 [H][H], fc2b [H][H], fc3 [dv][H]
struct McGrad { float* W[4]; float* b[4]; };

__device__ __forceinline__ uint32_t mc_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mc_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = mc_tf32(v);
    lo = mc_tf32(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mc_mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float w0, float w1) {
    uint32_t bh0, bl0, bh1, bl1;
    mc_split(w0, bh0, bl0);
    mc_split(w1, bh1, bl1);
    mc_mma(c, al, bh0, bh1);
    mc_mma(c, ah, bl0, bl1);
    mc_mma(c, ah, bh0, bh1);
}

constexpr int kMcNc = 2;   // n-tiles per warp pass

// acc[j] += in[16][K] * Bm for the n-tiles nt0 .. nt0 + kMcNc - 1.  `in`: the CTA's shared rows (pitch PI, pitch mod 8 = 4:
// conflict-free A fragments; columns K .. pad8(K) hold zeros).  WT = false: Bm[k][n] = W[n * ldw + k] (forward: W is [N][K],
// PyTorch layout); WT = true: Bm[k][n] = W[k * ldw + n] (input gradient: the same W read as [K][N]).  The weights of the next
// k-step are fetched (L1 / L2) before the current one is multiplied.
template <bool WT>
__device__ __forceinline__ void mc_wload(const float* __restrict__ W, int ldw, int K, int N, int nt0, int k0, int g, int t, float (&w)[2 * kMcNc]) {
    const int ka = k0 + t, kb = ka + 4;
#pragma unroll
    for (int j = 0; j < kMcNc; ++j) {
        const int n = 8 * (nt0 + j) + g;
        w[2 * j] = (n < N && ka < K) ? __ldg(WT ? W + (size_t)ka * ldw + n : W + (size_t)n * ldw + ka) : 0.f;
        w[2 * j + 1] = (n < N && kb < K) ? __ldg(WT ? W + (size_t)kb * ldw + n : W + (size_t)n * ldw + kb) : 0.f;
    }
}
template <bool WT>
__device__ __forceinline__ void mc_gemm(const float* in, int PI, int K, const float* __restrict__ W, int ldw, int N, int nt0, int lane,
                                        float (&acc)[kMcNc][4]) {
    const int g = lane >> 2, t = lane & 3, K8 = (K + 7) & ~7;
    float wc[2 * kMcNc], wn[2 * kMcNc];
    mc_wload<WT>(W, ldw, K, N, nt0, 0, g, t, wc);
    for (int k0 = 0; k0 < K8; k0 += 8) {
        if (k0 + 8 < K8) mc_wload<WT>(W, ldw, K, N, nt0, k0 + 8, g, t, wn);
        uint32_t ah[4], al[4];
        const float* ap = in + g * PI + k0 + t;
        mc_split(ap[0], ah[0], al[0]);
        mc_split(ap[8 * PI], ah[1], al[1]);
        mc_split(ap[4], ah[2], al[2]);
        mc_split(ap[8 * PI + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < kMcNc; ++j) mc_mma3(acc[j], ah, al, wc[2 * j], wc[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 2 * kMcNc; ++j) wc[j] = wn[j];
    }
}

enum { MC_EPI_RELU = 0, MC_EPI_BIAS = 1, MC_EPI_MASK = 2, MC_EPI_PLAIN = 3, MC_EPI_ADD = 4 };

// out[16][pad8(N)] = epilogue(in * Bm):  RELU: relu(. + b)   BIAS: . + b   MASK: . where mask > 0 else 0   PLAIN: .   ADD: out += .
// Columns N .. pad8(N) of out receive zeros (they are the next layer's K padding).  The CTA's warps take the n-tile pairs in turn;
// ends with a CTA barrier.
template <bool WT, int EPI>
__device__ __forceinline__ void mc_layer(const float* in, int PI, int K, const float* __restrict__ W, int ldw, int N, const float* __restrict__ b,
                                         const float* mask, float* out, int PO) {
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, NT = (N + 7) >> 3;
    for (int nt0 = kMcNc * warp; nt0 < NT; nt0 += kMcNc * nw) {
        float acc[kMcNc][4] = {};
        mc_gemm<WT>(in, PI, K, W, ldw, N, nt0, lane, acc);
#pragma unroll
        for (int j = 0; j < kMcNc; ++j) {
            const int c = 8 * (nt0 + j) + 2 * t;
            if (nt0 + j < NT) {
                float v[4] = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
                float* o0 = out + g * PO + c;
                float* o1 = o0 + 8 * PO;
                if (EPI == MC_EPI_RELU || EPI == MC_EPI_BIAS) {
                    const float b0 = c < N ? __ldg(b + c) : 0.f, b1 = c + 1 < N ? __ldg(b + c + 1) : 0.f;
                    v[0] += b0; v[1] += b1; v[2] += b0; v[3] += b1;
                    if (EPI == MC_EPI_RELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
                    }
                }
                if (EPI == MC_EPI_MASK) {
                    const float* m0 = mask + g * PO + c;
                    const float* m1 = m0 + 8 * PO;
                    v[0] = m0[0] > 0.f ? v[0] : 0.f; v[1] = m0[1] > 0.f ? v[1] : 0.f;
                    v[2] = m1[0] > 0.f ? v[2] : 0.f; v[3] = m1[1] > 0.f ? v[3] : 0.f;
                }
                if (EPI == MC_EPI_ADD) { v[0] += o0[0]; v[1] += o0[1]; v[2] += o1[0]; v[3] += o1[1]; }
                *reinterpret_cast<float2*>(o0) = make_float2(v[0], v[1]);
                *reinterpret_cast<float2*>(o1) = make_float2(v[2], v[3]);
            }
        }
    }
    __syncthreads();
}

// global rows [nrows][n] (contiguous) <-> the CTA's shared rows of pitch P; rows beyond nrows read as zeros
__device__ __forceinline__ void mc_load_rows(float* dst, int P, const float* __restrict__ src, int n, int nrows) {
    const float inv = 1.f / (float)n;
    for (int i = threadIdx.x; i < 16 * n; i += blockDim.x) {
        const int r = (int)(((float)i + 0.5f) * inv);
        dst[r * P + (i - r * n)] = r < nrows ? __ldg(src + i) : 0.f;
    }
}
__device__ __forceinline__ void mc_store_rows(const float* src, int P, float* __restrict__ dst, int n, int nrows) {
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int nq = n >> 2;
        const float inv = 1.f / (float)nq;
        for (int i = threadIdx.x; i < nrows * nq; i += blockDim.x) {
            const int r = (int)(((float)i + 0.5f) * inv);
            reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<const float4*>(src + r * P + 4 * (i - r * nq));
        }
    } else {
        const float inv = 1.f / (float)n;
        for (int i = threadIdx.x; i < nrows * n; i += blockDim.x) {
            const int r = (int)(((float)i + 0.5f) * inv);
            dst[i] = src[r * P + (i - r * n)];
        }
    }
}

struct McDims {
    int du, dv, H, PU, P, PS;
    __host__ __device__ McDims(int du_, int dv_, int H_) : du(du_), dv(dv_), H(H_), PU(((du_ + 7) & ~7) + 4), P(((H_ + 7) & ~7) + 4), PS(((dv_ + 7) & ~7) + 4) {}
    __host__ __device__ int fwd_floats() const { return 16 * (PU + 2 * P + PS) + 16 * (kMcThreads / 32); }
    __host__ __device__ int bwd_floats() const { return 16 * (2 * PU + 5 * P + 2 * PS); }
};

struct McFwdArgs {
    McNet net[2];          // s, t
    const float* u; const float* v; float* y; float* jac;
    long long B; int du, dv, H; float alpha; int rev;
};

// One CTA = one tile of 16 samples at a time; its warps split every layer's output columns.  (At most 8 warps are launched; the
// register cap of 3 CTAs x 256 threads leaves the forward at 76 registers without spills and five CTAs of 5 warps per SM for the
// lens shape: hidden = 224 forward 330 -> 265 us.  The backward kernel spills under the same cap and is slower: left alone.)
__global__ void __launch_bounds__(256, 3) mc_forward_kernel(McFwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const McDims D(a.du, a.dv, a.H);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* U = sm;
    float* A = U + 16 * D.PU;
    float* Bf = A + 16 * D.P;
    float* S = Bf + 16 * D.P;
    float* JP = S + 16 * D.PS;                         // [warp][16] log-det partials
    for (int i = threadIdx.x; i < D.fwd_floats(); i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long ntiles = (a.B + 15) / 16;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * 16;
        const int nrows = (int)(a.B - r0 < 16 ? a.B - r0 : 16);
        mc_load_rows(U, D.PU, a.u + r0 * a.du, a.du, nrows);
        __syncthreads();
        for (int net = 0; net < 2; ++net) {
            const McNet& w = a.net[net];
            mc_layer<false, MC_EPI_RELU>(U, D.PU, a.du, w.W[0], a.du, a.H, w.b[0], nullptr, A, D.P);
            mc_layer<false, MC_EPI_RELU>(A, D.P, a.H, w.W[1], a.H, a.H, w.b[1], nullptr, Bf, D.P);
            mc_layer<false, MC_EPI_RELU>(Bf, D.P, a.H, w.W[2], a.H, a.H, w.b[2], nullptr, A, D.P);
            if (net == 0) {
                mc_layer<false, MC_EPI_BIAS>(A, D.P, a.H, w.W[3], a.H, a.dv, w.b[3], nullptr, S, D.PS);
            } else {
                // t arrives in the C fragments: the coupling runs in the epilogue
                float jr[2] = {0.f, 0.f};
                const int NT = (a.dv + 7) >> 3;
                for (int nt0 = kMcNc * warp; nt0 < NT; nt0 += kMcNc * nw) {
                    float acc[kMcNc][4] = {};
                    mc_gemm<false>(A, D.P, a.H, w.W[3], a.H, a.dv, nt0, lane, acc);
#pragma unroll
                    for (int j = 0; j < kMcNc; ++j) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 8 * (nt0 + j) + 2 * t + (e & 1), rr = g + 8 * (e >> 1);
                            if (nt0 + j < NT && c < a.dv && rr < nrows) {
                                const float tt = acc[j][e] + __ldg(w.b[3] + c);
                                const float la = a.alpha * atanf(S[rr * D.PS + c]);
                                const float vv = __ldg(a.v + (r0 + rr) * a.dv + c);
                                a.y[(r0 + rr) * a.dv + c] = a.rev ? (vv - tt) * expf(-la) : fmaf(expf(la), vv, tt);
                                jr[e >> 1] += la;
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    jr[e] += __shfl_xor_sync(0xffffffffu, jr[e], 1);
                    jr[e] += __shfl_xor_sync(0xffffffffu, jr[e], 2);
                    if (t == 0) JP[warp * 16 + g + 8 * e] = jr[e];
                }
                __syncthreads();
                if ((int)threadIdx.x < nrows) {
                    float j = 0.f;
                    for (int ww = 0; ww < nw; ++ww) j += JP[ww * 16 + threadIdx.x];   // fixed order
                    a.jac[r0 + threadIdx.x] = a.rev ? -j : j;
                }
            }
        }
    }
}

struct McBwdArgs {
    McNet net[2];
    const float* u; const float* v; const float* dy; const float* djac;    // djac may be null (= 0)
    float* du_grad; float* dv_grad;
    float* act;            // workspace: per net H1, H2, H3, G1, G2, G3 [B][H] and D3 [B][dv]
    long long B; int du, dv, H; float alpha;
};
__host__ __device__ inline size_t mc_net_floats(long long B, int dv, int H) { return (size_t)B * (6 * (size_t)H + dv); }

// rev = 0 direction only (the training direction, train_conditional.py:119-156)
__global__ void __launch_bounds__(320, 2) mc_backward_kernel(McBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const McDims D(a.du, a.dv, a.H);
    float* U = sm;
    float* DU = U + 16 * D.PU;
    float* H1 = DU + 16 * D.PU;
    float* H2 = H1 + 16 * D.P;
    float* H3 = H2 + 16 * D.P;
    float* Ga = H3 + 16 * D.P;
    float* Gb = Ga + 16 * D.P;
    float* S = Gb + 16 * D.P;
    float* D3 = S + 16 * D.PS;
    for (int i = threadIdx.x; i < D.bwd_floats(); i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long ntiles = (a.B + 15) / 16;
    const size_t BH = (size_t)a.B * a.H;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * 16;
        const int nrows = (int)(a.B - r0 < 16 ? a.B - r0 : 16);
        mc_load_rows(U, D.PU, a.u + r0 * a.du, a.du, nrows);
        __syncthreads();
        for (int net = 0; net < 2; ++net) {
            const McNet& w = a.net[net];
            float* base = a.act + net * mc_net_floats(a.B, a.dv, a.H);
            mc_layer<false, MC_EPI_RELU>(U, D.PU, a.du, w.W[0], a.du, a.H, w.b[0], nullptr, H1, D.P);
            mc_layer<false, MC_EPI_RELU>(H1, D.P, a.H, w.W[1], a.H, a.H, w.b[1], nullptr, H2, D.P);
            mc_layer<false, MC_EPI_RELU>(H2, D.P, a.H, w.W[2], a.H, a.H, w.b[2], nullptr, H3, D.P);
            if (net == 0) mc_layer<false, MC_EPI_BIAS>(H3, D.P, a.H, w.W[3], a.H, a.dv, w.b[3], nullptr, S, D.PS);
            // gradient at the subnet output: ds = (dy v e(s) + dlogdet) alpha / (1 + s^2), dt = dy; and dL/dv = dy e(s)
            {
                const float inv = 1.f / (float)a.dv;
                for (int i = threadIdx.x; i < 16 * a.dv; i += blockDim.x) {
                    const int r = (int)(((float)i + 0.5f) * inv), c = i - r * a.dv;
                    float d3 = 0.f;
                    if (r < nrows) {
                        const float dyv = __ldg(a.dy + (r0 + r) * a.dv + c);
                        if (net == 0) {
                            const float s = S[r * D.PS + c], e = expf(a.alpha * atanf(s));
                            const float dj = a.djac ? __ldg(a.djac + r0 + r) : 0.f;
                            d3 = fmaf(dyv * __ldg(a.v + (r0 + r) * a.dv + c), e, dj) * a.alpha / fmaf(s, s, 1.f);
                            a.dv_grad[(r0 + r) * a.dv + c] = dyv * e;
                        } else {
                            d3 = dyv;
                        }
                    }
                    D3[r * D.PS + c] = d3;
                }
            }
            mc_store_rows(H1, D.P, base + 0 * BH + r0 * a.H, a.H, nrows);
            mc_store_rows(H2, D.P, base + 1 * BH + r0 * a.H, a.H, nrows);
            mc_store_rows(H3, D.P, base + 2 * BH + r0 * a.H, a.H, nrows);
            __syncthreads();
            mc_store_rows(D3, D.PS, base + 6 * BH + r0 * a.dv, a.dv, nrows);
            // masked input-gradient chain; H3 / H2 / H1 double as the relu masks
            mc_layer<true, MC_EPI_MASK>(D3, D.PS, a.dv, w.W[3], a.H, a.H, nullptr, H3, Ga, D.P);
            mc_store_rows(Ga, D.P, base + 5 * BH + r0 * a.H, a.H, nrows);
            mc_layer<true, MC_EPI_MASK>(Ga, D.P, a.H, w.W[2], a.H, a.H, nullptr, H2, Gb, D.P);
            mc_store_rows(Gb, D.P, base + 4 * BH + r0 * a.H, a.H, nrows);
            mc_layer<true, MC_EPI_MASK>(Gb, D.P, a.H, w.W[1], a.H, a.H, nullptr, H1, Ga, D.P);
            mc_store_rows(Ga, D.P, base + 3 * BH + r0 * a.H, a.H, nrows);
            if (net == 0) mc_layer<true, MC_EPI_PLAIN>(Ga, D.P, a.H, w.W[0], a.du, a.du, nullptr, nullptr, DU, D.PU);
            else mc_layer<true, MC_EPI_ADD>(Ga, D.P, a.H, w.W[0], a.du, a.du, nullptr, nullptr, DU, D.PU);
        }
        mc_store_rows(DU, D.PU, a.du_grad + r0 * a.du, a.du, nrows);
        __syncthreads();
    }
}

// dW[n][k] = sum_s Dm[s][n] X[s][k], db[n] = sum_s Dm[s][n] over this block's share of the samples (gridDim.y splits); with more
// than one split the results go to partial buffers that mc_wgrad_reduce_kernel sums in a fixed order.
struct McWgProb { const float* Dm; const float* X; float* dW; float* db; int N, K, tiles_k, first_block; long long poff; };
struct McWgArgs { McWgProb p[8]; long long B; float* partial; long long pstride; };

__device__ __forceinline__ void mc_wg_load(const McWgProb& p, long long s0, long long s_end, int n0, int k0, int g, int t, float (&av)[4], float (&xv)[8]) {
    const bool sa = s0 + t < s_end, sb = s0 + t + 4 < s_end;
    const bool na = n0 + g < p.N, nb = n0 + g + 8 < p.N;
    const float* da = p.Dm + (s0 + t) * p.N + n0 + g;
    const float* db_ = da + 4 * (size_t)p.N;
    av[0] = sa && na ? __ldg(da) : 0.f; av[1] = sa && nb ? __ldg(da + 8) : 0.f;
    av[2] = sb && na ? __ldg(db_) : 0.f; av[3] = sb && nb ? __ldg(db_ + 8) : 0.f;
    const float* xa = p.X + (s0 + t) * p.K + k0 + g;
    const float* xb = xa + 4 * (size_t)p.K;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool kin = k0 + 8 * j + g < p.K;
        xv[2 * j] = sa && kin ? __ldg(xa + 8 * j) : 0.f;
        xv[2 * j + 1] = sb && kin ? __ldg(xb + 8 * j) : 0.f;
    }
}

__global__ void __launch_bounds__(32 * kMcWgWarps) mc_wgrad_kernel(McWgArgs a) {
    __shared__ float red[kMcWgWarps][32][17];
    __shared__ float redb[kMcWgWarps][16];
    int pi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) pi += (int)blockIdx.x >= a.p[i].first_block;
    const McWgProb& p = a.p[pi];
    const int tile = blockIdx.x - p.first_block, tn = tile / p.tiles_k, tk = tile - tn * p.tiles_k, n0 = 16 * tn, k0 = 32 * tk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // this block's samples
    const long long per = ((a.B + gridDim.y - 1) / gridDim.y + 7) & ~7LL;
    const long long s_begin = per * blockIdx.y, s_end = s_begin + per < a.B ? s_begin + per : a.B;
    float acc[4][4] = {};
    float bs[2] = {0.f, 0.f};
    float av[4], xv[8], avn[4], xvn[8];
    long long s0 = s_begin + 8 * warp;
    if (s0 < s_end) mc_wg_load(p, s0, s_end, n0, k0, g, t, av, xv);
    for (; s0 < s_end; s0 += 8 * kMcWgWarps) {
        if (s0 + 8 * kMcWgWarps < s_end) mc_wg_load(p, s0 + 8 * kMcWgWarps, s_end, n0, k0, g, t, avn, xvn);
        bs[0] += av[0] + av[2]; bs[1] += av[1] + av[3];
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mc_split(av[i], ah[i], al[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + 8 * j < p.K) mc_mma3(acc[j], ah, al, xv[2 * j], xv[2 * j + 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = avn[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = xvn[i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) red[warp][lane][4 * j + e] = acc[j][e];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        bs[e] += __shfl_xor_sync(0xffffffffu, bs[e], 1);
        bs[e] += __shfl_xor_sync(0xffffffffu, bs[e], 2);
        if (t == 0) redb[warp][g + 8 * e] = bs[e];
    }
    __syncthreads();
    float* dW = gridDim.y > 1 ? a.partial + blockIdx.y * a.pstride + p.poff : p.dW;
    float* db = gridDim.y > 1 ? dW + (size_t)p.N * p.K : p.db;
    for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) {
        const int ln = i >> 4, q = i & 15, j = q >> 2, e = q & 3;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kMcWgWarps; ++w) v += red[w][ln][q];
        const int n = n0 + (ln >> 2) + 8 * (e >> 1), k = k0 + 8 * j + 2 * (ln & 3) + (e & 1);
        if (n < p.N && k < p.K) dW[(size_t)n * p.K + k] = v;
    }
    if (tk == 0 && threadIdx.x < 16 && n0 + (int)threadIdx.x < p.N) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kMcWgWarps; ++w) v += redb[w][threadIdx.x];
        db[n0 + threadIdx.x] = v;
    }
}

// out = sum over the splits' partial buffers (fixed order); partial layout per problem: dW [N][K] then db [N]
__global__ void mc_wgrad_reduce_kernel(McWgArgs a, int nsplit) {
    const McWgProb& p = a.p[blockIdx.y];
    const long long n = (long long)p.N * p.K + p.N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < nsplit; ++s) v += a.partial[s * a.pstride + p.poff + i];
        if (i < (long long)p.N * p.K) p.dW[i] = v;
        else p.db[i - (long long)p.N * p.K] = v;
    }
}

int mc_sms() {
    static const int sms = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
        return n;
    }();
    return sms > 0 ? sms : 148;
}

void mc_fill(McNet (&net)[2], const float* const* params) {
    for (int n = 0; n < 2; ++n)
        for (int l = 0; l < 4; ++l) { net[n].W[l] = params[8 * n + 2 * l]; net[n].b[l] = params[8 * n + 2 * l + 1]; }
}

// CTAs per SM of (kernel, block size, dynamic shared memory): the attribute call and the occupancy query cost ~10 us of host time,
// which matters for kernels that run for 20 us - asked once per distinct configuration
cudaError_t mc_occupancy(const void* fn, int threads, size_t smem, int* per_sm) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, size_t>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(fn, threads, smem);
    const auto it = cache.find(key);
    if (it != cache.end()) { *per_sm = it->second; return cudaSuccess; }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fn, threads, smem);
    if (e != cudaSuccess) return e;
    if (*per_sm < 1) return cudaErrorInvalidConfiguration;
    cache[key] = *per_sm;
    return cudaSuccess;
}

// warps per CTA (they split a layer's n-tile pairs) and the grid (one 16-sample tile per CTA at a time)
cudaError_t mc_geometry(const void* fn, int floats, int H, long long B, int* nw, size_t* smem, int* grid) {
    const int pairs = (((H + 7) >> 3) + kMcNc - 1) / kMcNc;
    *nw = pairs < 1 ? 1 : pairs > 8 ? 8 : pairs;
    *smem = sizeof(float) * (size_t)floats;
    if (*smem > 227 * 1024) return cudaErrorInvalidValue;
    int per_sm = 1;
    cudaError_t e = mc_occupancy(fn, 32 * *nw, *smem, &per_sm);
    if (e != cudaSuccess) return e;
    const long long ntiles = (B + 15) / 16, cap = (long long)mc_sms() * per_sm;
    *grid = (int)(ntiles < cap ? ntiles : cap);
    return cudaSuccess;
}

constexpr int kMcMaxSplit = 32;
// CTAs of the weight-gradient kernel that are resident at once on the device: the sample splits fill exactly one wave
int mc_wgrad_resident() {
    static const int n = [] {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)mc_wgrad_kernel, 32 * kMcWgWarps, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        return per_sm * mc_sms();
    }();
    return n;
}
size_t mc_param_floats(int du, int dv, int H) { return 2 * ((size_t)H * du + 2 * (size_t)H * H + (size_t)dv * H + 3 * (size_t)H + dv); }

}  // namespace

bool mc_supported(int du, int dv, int H) { return du >= 1 && du <= kMcMaxIO && dv >= 1 && dv <= kMcMaxIO && H >= 1 && H <= kMcMaxH; }

cudaError_t mc_forward(const float* u, int du, const float* v, int dv, int H, const float* const* params, float clamp, int rev, long long B,
                       float* y, float* logdet, cudaStream_t st) {
    if (!mc_supported(du, dv, H) || B < 0) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    McFwdArgs a;
    mc_fill(a.net, params);
    a.u = u; a.v = v; a.y = y; a.jac = logdet; a.B = B; a.du = du; a.dv = dv; a.H = H; a.alpha = clamp * 0.636f; a.rev = rev;
    int nw = 0, grid = 0; size_t smem = 0;
    cudaError_t e = mc_geometry((const void*)mc_forward_kernel, McDims(du, dv, H).fwd_floats(), H, B, &nw, &smem, &grid);
    if (e != cudaSuccess) return e;
    mc_forward_kernel<<<grid, 32 * nw, smem, st>>>(a); HINT_LAUNCHED();
    return cudaGetLastError();
}

size_t mc_workspace_bytes(int du, int dv, int H, long long B) {
    return sizeof(float) * (2 * mc_net_floats(B, dv, H) + kMcMaxSplit * mc_param_floats(du, dv, H)) + 256;
}

cudaError_t mc_backward(const float* u, int du, const float* v, int dv, int H, const float* const* params, float clamp, long long B,
                        const float* dy, const float* dlogdet, float* du_grad, float* dv_grad, float* const* dparams, void* ws,
                        size_t ws_bytes, cudaStream_t st) {
    if (!mc_supported(du, dv, H) || B < 0) return cudaErrorInvalidValue;
    if (ws_bytes < mc_workspace_bytes(du, dv, H, B)) return cudaErrorInvalidValue;
    if (B == 0) {
        for (int n = 0; n < 2; ++n)
            for (int l = 0; l < 4; ++l) {
                const size_t rows = l == 3 ? dv : H, cols = l == 0 ? du : H;
                cudaError_t e = cudaMemsetAsync(dparams[8 * n + 2 * l], 0, sizeof(float) * rows * cols, st);
                if (e == cudaSuccess) e = cudaMemsetAsync(dparams[8 * n + 2 * l + 1], 0, sizeof(float) * rows, st);
                if (e != cudaSuccess) return e;
            }
        return cudaSuccess;
    }
    McBwdArgs a;
    mc_fill(a.net, params);
    a.u = u; a.v = v; a.dy = dy; a.djac = dlogdet; a.du_grad = du_grad; a.dv_grad = dv_grad; a.act = static_cast<float*>(ws);
    a.B = B; a.du = du; a.dv = dv; a.H = H; a.alpha = clamp * 0.636f;
    int nw = 0, grid = 0; size_t smem = 0;
    cudaError_t e = mc_geometry((const void*)mc_backward_kernel, McDims(du, dv, H).bwd_floats(), H, B, &nw, &smem, &grid);
    if (e != cudaSuccess) return e;
    mc_backward_kernel<<<grid, 32 * nw, smem, st>>>(a); HINT_LAUNCHED();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    McWgArgs wg;
    wg.B = B;
    wg.partial = a.act + 2 * mc_net_floats(B, dv, H);
    wg.pstride = (long long)mc_param_floats(du, dv, H);
    long long poff = 0;
    int blocks = 0;
    const size_t BH = (size_t)B * H;
    for (int n = 0; n < 2; ++n) {
        const float* base = a.act + n * mc_net_floats(B, dv, H);
        for (int l = 0; l < 4; ++l) {
            McWgProb& p = wg.p[4 * n + l];
            p.N = l == 3 ? dv : H;
            p.K = l == 0 ? du : H;
            p.Dm = l == 3 ? base + 6 * BH : base + (3 + l) * BH;          // G1, G2, G3, D3
            p.X = l == 0 ? u : base + (l - 1) * BH;                        // u, H1, H2, H3
            p.dW = dparams[8 * n + 2 * l];
            p.db = dparams[8 * n + 2 * l + 1];
            p.tiles_k = (p.K + 31) / 32;
            p.first_block = blocks;
            p.poff = poff;
            poff += (long long)p.N * p.K + p.N;
            blocks += ((p.N + 15) / 16) * p.tiles_k;
        }
    }
    // one full wave of CTAs: split the samples when the output tiles alone are too few (each split >= 64 samples)
    long long nsplit = mc_wgrad_resident() / blocks;
    const long long by_rows = (B + 63) / 64;
    nsplit = nsplit > by_rows ? by_rows : nsplit;
    nsplit = nsplit > kMcMaxSplit ? kMcMaxSplit : nsplit < 1 ? 1 : nsplit;
    mc_wgrad_kernel<<<dim3(blocks, (unsigned)nsplit), 32 * kMcWgWarps, 0, st>>>(wg); HINT_LAUNCHED();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (nsplit > 1) {
        mc_wgrad_reduce_kernel<<<dim3(32, 8), 256, 0, st>>>(wg, (int)nsplit); HINT_LAUNCHED();
    }
    return cudaGetLastError();
}

// out[N][K] = Dm^T X for Dm [B][N], X [B][K] with the weight-gradient kernel above (one problem, more sample splits): used by the
// Householder mixing's dW = x^T dy.  Workspace: N floats (unused bias sums) + splits x (N K + N) partials.
constexpr int kMcXtySplit = 128;
size_t mc_xt_y_workspace_bytes(int N, int K) { return sizeof(float) * ((size_t)N + (size_t)kMcXtySplit * ((size_t)N * K + N)) + 64; }

cudaError_t mc_xt_y(const float* Dm, const float* X, int N, int K, long long B, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (N < 1 || K < 1 || B < 0 || ws_bytes < mc_xt_y_workspace_bytes(N, K)) return cudaErrorInvalidValue;
    if (B == 0) return cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * K, st);
    McWgArgs wg;
    wg.B = B;
    float* f = static_cast<float*>(ws);
    wg.partial = f + N;
    wg.pstride = (long long)N * K + N;
    for (int i = 0; i < 8; ++i) { wg.p[i] = McWgProb{}; wg.p[i].first_block = 0x7fffffff; }
    McWgProb& p = wg.p[0];
    p.Dm = Dm; p.X = X; p.dW = out; p.db = f; p.N = N; p.K = K; p.tiles_k = (K + 31) / 32; p.first_block = 0; p.poff = 0;
    const int blocks = ((N + 15) / 16) * p.tiles_k;
    long long nsplit = mc_wgrad_resident() / blocks;
    const long long by_rows = (B + 255) / 256;
    nsplit = nsplit > by_rows ? by_rows : nsplit;
    nsplit = nsplit > kMcXtySplit ? kMcXtySplit : nsplit < 1 ? 1 : nsplit;
    mc_wgrad_kernel<<<dim3(blocks, (unsigned)nsplit), 32 * kMcWgWarps, 0, st>>>(wg); HINT_LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (nsplit > 1) {
        mc_wgrad_reduce_kernel<<<dim3(32, 1), 256, 0, st>>>(wg, (int)nsplit); HINT_LAUNCHED();
    }
    return cudaGetLastError();
}

}  // namespace hint
