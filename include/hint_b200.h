/*
 * hint_b200 — C ABI of the B200-native HINT coupling-block hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: the
 * hot path is the pure-Python module /root/reference/hint.py.  Each entry point below replaces the
 * part of that file named in its comment; the Python host (hint_b200/block.py) binds these symbols
 * with ctypes and keeps hint.py's nn.Module surface on top of them.
 *
 * Conventions
 *   - All tensors are dense fp32, row-major, device pointers on the CURRENT CUDA device, base
 *     pointers 16-byte aligned.  x/z/dz/dx are [B, d]; c/dc are [B, dc] (all conditions of
 *     hint.py:76 concatenated in list order; NULL when dc == 0); logdet/dlogdet are [B].
 *   - `params` / `dparams` are flat vectors in the reference's `parameters()` order: pre-order over
 *     tree nodes, per node s.0.weight, s.0.bias, s.2.weight, s.2.bias, s.4.weight, s.4.bias, then
 *     t.* (hint.py:44-45,49-52; nn.Linear weights are [out, in]).  hint_plan_param_layout() gives
 *     the offsets, so a host can expose reference-named views of one flat buffer.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); no host sync inside.
 *   - Return value: HINT_OK or an error code; hint_last_error() returns a thread-local message.
 *   - There is no CPU path: calling a compute entry point without a usable sm_100 device fails.
 */
#ifndef HINT_B200_H
#define HINT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HINT_OK 0
#define HINT_ERR_INVALID 1      /* bad shape / NULL pointer / misaligned pointer               */
#define HINT_ERR_UNSUPPORTED 2  /* option outside the fused path (conv, custom subnet, reshuffle) */
#define HINT_ERR_CUDA 3         /* CUDA runtime error (message carries cudaGetErrorString)      */
#define HINT_ERR_WORKSPACE 4    /* workspace smaller than hint_workspace_bytes()                */

/* arithmetic mode of the subnet GEMMs (accumulation is always fp32) */
#define HINT_MODE_FP32 0        /* CUDA-core FFMA, bit-for-bit fp32 products                    */
#define HINT_MODE_TF32 1        /* tensor cores, operands rounded to 10-bit mantissa (warp-MMA fused-tree kernels;
                                   forward, inverse and backward)                                 */
#define HINT_MODE_TF32X3 2      /* same kernels, 3xTF32 split (big*big + small*big + big*small): fp32-class accuracy */
#define HINT_MODE_TF32_TCGEN05 3 /* alias of HINT_MODE_TF32_TC3 (one tcgen05 / TMEM kernel runs forward, inverse and backward)  */
#define HINT_MODE_TF32_MMA 4    /* HINT_MODE_TF32 with the warp-MMA kernel forced for forward/inverse too (HINT_MODE_TF32
                                   itself picks the faster of the two forward kernels the block fits)                */

#define HINT_MODE_TF32_CHAIN 5  /* HINT_MODE_TF32 with the register-chained warp-MMA kernels forced (HINT_MODE_TF32 picks
                                   them whenever the block fits their shape table)                                   */

#define HINT_MODE_TF32_TC3 6    /* HINT_MODE_TF32 with the tcgen05 / TMEM kernel forced: forward / inverse transport programs and the
                                   memory-free backward (tensor-memory accumulators, weight gradients as tcgen05.mma over shared-
                                   memory images); HINT_MODE_TF32 picks it for the blocks the register-chained kernels do not cover */

/* which workspace hint_workspace_bytes() sizes */
#define HINT_WS_FORWARD 0
#define HINT_WS_BACKWARD 1

typedef struct hint_plan hint_plan_t;

/* One tree node of hint.py:25-54, flattened.  Nodes are numbered in pre-order (root 0, then the
 * whole upper subtree, then the lower subtree); every node owns the contiguous input columns
 * [lo, hi).  upper half = [lo, lo+k), lower half = [lo+k, hi)  (hint.py:41,68). */
typedef struct hint_node_info {
    int32_t depth, lo, hi, k;
    int32_t cin;    /* k + dc             (hint.py:44) */
    int32_t h;      /* hidden width       (hint.py:31-34,50) */
    int32_t cout;   /* (hi - lo) - k      (hint.py:44) */
    int32_t leaf;   /* hint.py:47,54 */
    int32_t parent, upper, lower;
    int64_t param_offset; /* offset of this node's s.0.weight in the flat parameter vector */
} hint_node_info_t;

/* --- plan: replaces HierarchicalAffineCouplingTree.__init__ (hint.py:25-54) and the ctor argument
 * checks of HierarchicalAffineCouplingBlock.__init__ (hint.py:108-122).
 * c_internal/n_internal: hidden widths per depth (empty -> [d], last entry repeats).
 * reshuffle != 0 is rejected with HINT_ERR_UNSUPPORTED at this level: the per-node mixings of hint.py:36-39,64-65,93-94 compose
 * into ONE d x d orthogonal matrix in front of the un-shuffled tree, which the host applies with hint_householder_apply (see
 * hint_b200/block.py); the plan itself is always the un-shuffled tree. */
int hint_plan_create(int32_t d, int32_t dc, const int32_t* c_internal, int32_t n_internal, double clamp,
                     int32_t max_splits, int32_t min_split_size, int32_t reshuffle, hint_plan_t** out);
void hint_plan_destroy(hint_plan_t* plan);

int32_t hint_plan_num_nodes(const hint_plan_t* plan);
int hint_plan_node(const hint_plan_t* plan, int32_t idx, hint_node_info_t* out);
int64_t hint_plan_param_count(const hint_plan_t* plan);
/* offsets[n*12 + net*6 + layer*2 + kind]: net 0=s 1=t, layer 0..2, kind 0=weight 1=bias */
int hint_plan_param_layout(const hint_plan_t* plan, int64_t* offsets, int64_t n_offsets);
/* algorithmic forward FLOPs per sample: sum_nodes 2*2*(cin*h + h*h + h*cout)   */
int64_t hint_plan_flops_per_sample(const hint_plan_t* plan);
/* samples per CTA tile chosen for the forward / backward schedule (for reporting) */
int32_t hint_plan_tile_rows(const hint_plan_t* plan, int32_t which);

/* 1 when `mode` can run this block (every kernel family has a shared-memory / TMEM envelope; FP32 is the widest) */
int32_t hint_plan_mode_supported(const hint_plan_t* plan, int32_t mode);

size_t hint_workspace_bytes(const hint_plan_t* plan, int64_t B, int32_t which);

/* --- forward / inverse transport + log|det J|: replaces HierarchicalAffineCouplingTree.forward
 * (hint.py:62-101) as called by HierarchicalAffineCouplingBlock.forward (hint.py:124-126).
 * rev == 0: z = f(x; c), logdet = +sum log e(s).   rev != 0: z = f^-1(x; c), logdet = -sum log e(s).
 * x is not modified; z must not alias x. */
int hint_forward(const hint_plan_t* plan, const float* x, const float* c, const float* params, int64_t B,
                 int32_t rev, int32_t mode, float* z, float* logdet, void* workspace, size_t workspace_bytes,
                 void* stream);

/* --- backward of the rev == 0 direction: replaces the autograd tape the reference builds over
 * hint.py:62-101 (triggered at train_unconditional.py:137).  Memory-free: takes the block OUTPUT z
 * (and c), re-derives every node's input by running the inverse sweep inside the kernel, and emits
 * dx [B,d], dc [B,dc] (NULL allowed), dparams [param_count] (overwritten, not accumulated) for
 * upstream gradients dz [B,d] and dlogdet [B].  x_rec (optional, may be NULL) receives the
 * reconstructed block input f^-1(z). */
int hint_backward(const hint_plan_t* plan, const float* z, const float* c, const float* params,
                  const float* dz, const float* dlogdet, int64_t B, int32_t mode, float* x_rec, float* dx,
                  float* dc, float* dparams, void* workspace, size_t workspace_bytes, void* stream);

/* --- training edge: the step around the blocks as this library's own kernels ------------------------
 * hint_backward_nll: hint_backward for the blocks of a flow trained with the reference's NLL loss
 * (train_unconditional.py:128-132: L = 0.5*mean_b|z_b|^2 - mean_b sum_blocks logdet_b).  dlogdet is never read from memory
 * (it is -grad_scale for every block, grad_scale = 1/B_global); for the LAST block pass dz = NULL: dz = grad_scale * z is
 * generated in the kernel's tile load as well.  Earlier blocks pass the dx of the block after them as dz.
 * Implemented by the register-chained and tcgen05 training kernels (HINT_MODE_TF32 / _CHAIN / _TC3); other modes return
 * HINT_ERR_UNSUPPORTED and the host materialises the gradients itself. */
int hint_backward_nll(const hint_plan_t* plan, const float* z, const float* c, const float* params, const float* dz,
                      float grad_scale, int64_t B, int32_t mode, float* x_rec, float* dx, float* dc, float* dparams,
                      void* workspace, size_t workspace_bytes, void* stream);
/* out[i] = x[i] + sigma * N(0,1), i < n: replaces `x += noise * torch.randn_like(x)` (train_unconditional.py:121-123).
 * Counter-based Philox4x32-10 + Box-Muller: (seed, offset) identify the stream, element i owns counter i/4; out may alias x. */
int hint_add_noise(const float* x, float* out, int64_t n, float sigma, uint64_t seed, uint64_t offset, void* stream);
/* loss3[0] = 0.5*mean_b|z_b|^2 - mean_b logdet_b, loss3[1], loss3[2] = the two terms (device floats); logdet = the sum of the
 * n_logdets (1..64) per-block [B] vectors whose device pointers the HOST array `logdets` holds (the per-block log-dets of a
 * flow are summed inside the reduction).  Deterministic two-stage fp64 reduction.  Replaces train_unconditional.py:128-132. */
size_t hint_nll_workspace_bytes(void);
int hint_nll_loss(const float* z, const float* const* logdets, int32_t n_logdets, int64_t B, int32_t d, float* loss3,
                  void* workspace, size_t workspace_bytes, void* stream);
/* One optimizer step over n_tensors flat fp32 tensors in ONE launch: g = clamp(grad, +-grad_clamp) (skipped when
 * grad_clamp <= 0), then torch.optim.Adam's update with L2 weight decay and bias correction for step number `step` >= 1.
 * Replaces train_unconditional.py:141-144 with the optimizer of :174-176.  The pointer arrays live on the host. */
int hint_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float grad_clamp, int64_t step, void* stream);

/* --- inter-block orthogonal mixing: FrEIA `HouseholderPerm` as the reference's configs use it between HINT
 * blocks (configs/uci_data/miniboone_hint_8.py:60-63 fixed, configs/plus_shape/unconditional_hint_4_3.py:60-71 trainable).
 * W [d,d] row-major = prod_{i < n_reflections} (I - 2 v_i v_i^T / |v_i|^2), v_i = Vs[i,:]; forward y = x W, reverse y = x W^T,
 * log|det| = 0.  FrEIA's source is not part of the reference: published definition, parity-unpinned.  1 <= d <= 128.
 *   hint_householder_matrix           W from Vs (one launch; run once per step when the reflections are trainable)
 *   hint_householder_matrix_backward  dVs [n_reflections,d] from dW and the W the forward produced (no stored intermediates)
 *   hint_householder_apply            y = x W (transpose = 0) or x W^T (transpose != 0), fp32-grade (3 x TF32 split products, fp32 accumulation; error <= 2 x an fp32 GEMM's); any 4-byte alignment; also the input gradient
 *                                     (dx = dy W^T)
 *   hint_householder_wgrad            dW = x^T dy (deterministic two-stage reduction) */
int hint_householder_matrix(const float* Vs, int32_t n_reflections, int32_t d, float* W, void* stream);
int hint_householder_matrix_backward(const float* Vs, const float* W, const float* dW, int32_t n_reflections, int32_t d,
                                     float* dVs, void* stream);
int hint_householder_apply(const float* x, const float* W, int64_t B, int32_t d, int32_t transpose, float* y, void* stream);
size_t hint_householder_wgrad_workspace_bytes(int32_t d);
int hint_householder_wgrad(const float* x, const float* dz, int64_t B, int32_t d, float* dW, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- baseline couplings of the 2-lane conditional configs ------------------------------------------------
 * FrEIA `AffineCoupling` / `ExternalAffineCoupling` with `F_fully_connected` subnets (fc1-ReLU-fc2-ReLU-fc2b-ReLU-fc3), as
 * configs/lens_shape/conditional_hint_8_full.py:78-89 places them either side of the HINT block.  FrEIA's source is not part of
 * the reference: published definition, parity-unpinned; the checker is the plain-PyTorch statement in FrEIA/modules/coupling.py.
 *     y = e(s(u)) * v + t(u),  logdet = sum_c log e(s_c),  e(s) = exp(clamp * 0.636 * atan(s))     (rev: y = (v - t(u)) / e(s(u)))
 * u [B,du]: the subnets' input (x1 | condition, or the condition alone); v [B,dv]: the transformed part; hidden: internal_size.
 * params / dparams: 16 device pointers = [s-net, t-net] x [fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc2b.weight, fc2b.bias,
 * fc3.weight, fc3.bias], weights in PyTorch's [out][in] row-major layout (the modules' own tensors: no repacking).
 * Envelope: du, dv <= 128, hidden <= 256 (hint_mlp_coupling_supported).  fp32-grade arithmetic (error-compensated 3 x TF32
 * tensor-core products, fp32 accumulation).  One launch forward; backward (rev = 0 direction) = one fused launch + one
 * weight-gradient launch (+ a fixed-order reduction over sample splits when the layers are small): deterministic, no atomics;
 * dlogdet may be NULL (= 0).  Pointers need 4-byte alignment only. */
int hint_mlp_coupling_supported(int32_t du, int32_t dv, int32_t hidden);
int hint_mlp_coupling_forward(const float* u, int32_t du, const float* v, int32_t dv, int32_t hidden, const float* const* params,
                              float clamp, int32_t rev, int64_t B, float* y, float* logdet, void* stream);
size_t hint_mlp_coupling_workspace_bytes(int32_t du, int32_t dv, int32_t hidden, int64_t B);
int hint_mlp_coupling_backward(const float* u, int32_t du, const float* v, int32_t dv, int32_t hidden, const float* const* params,
                               float clamp, int64_t B, const float* dy, const float* dlogdet, float* du_grad, float* dv_grad,
                               float* const* dparams, void* workspace, size_t workspace_bytes, void* stream);

/* ---- evaluation metric of the sampling scripts (rejection_sampling.py:56-73 `multi_mmd`) ---------------------
 *     out[0] = mean_ij [ k(|x_i - x_j|^2) + k(|y_i - y_j|^2) - 2 k(|x_i - y_j|^2) ],  k(D) = sum_w C_w^a_w ((C_w + D) / a_w)^(-a_w)
 * x, y [n,d] device, widths C_w / exponents a_w: HOST arrays of n_kernels <= 8 positive values (the script's default:
 * {0.5, 0.2, 0.2} / {1, 1, 0.5}); out: one device float.  One fused pair-tile kernel + a fixed-order fp64 reduction
 * (deterministic); nothing of size n x n is materialised. */
size_t hint_mmd_workspace_bytes(int64_t n);
int hint_multi_mmd(const float* x, const float* y, int64_t n, int32_t d, const float* widths, const float* exponents, int32_t n_kernels,
                   float* out, void* workspace, size_t workspace_bytes, void* stream);

const char* hint_last_error(void);
/* "hint_b200 <version> sm_100a" — lets the host check it loaded the in-tree build */
const char* hint_version(void);
/* number of CUDA kernels this library has launched in this process so far (all streams; monotone).  A host that wants to
 * report how many of the library's kernels ran inside a timed region takes the difference of two reads. */
uint64_t hint_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HINT_B200_H */
