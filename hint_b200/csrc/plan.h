// Host-side plan of one HINT coupling block: the recursion of hint.py:25-54 flattened into a node
// table, the canonical (reference `parameters()` order) and kernel-packed weight layouts, and the
// level-synchronous stage schedules the fused kernels walk.  Pure C++ (no CUDA) so it can be unit
// tested on a CPU-only machine.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hint_b200.h"

namespace hint {

constexpr int kThreads = 512;          // threads per CTA of the SIMT kernels
constexpr int kSmemMax = 227 * 1024;   // usable shared memory per CTA on sm_100

// ---- device-visible POD descriptors -------------------------------------------------------------
// One "column group": 4 consecutive output columns of one small GEMM  out[TM x 4] = in[TM x K] * W[K x 4].
// Activations live in shared memory column-major ([column][sample], column stride TM+4), so a
// column group reads K input columns and writes 4 output columns.
struct CG {
    int w_off;   // packed-weight offset of W[0][n0]; row k is at w_off + k*ldw
    int ldw;
    int b_off;   // packed offset of bias[n0..n0+3], or -1
    int in0, k0; // first input segment: smem column base, count
    int in1, k1; // second input segment (condition columns), count may be 0
    int out0;    // first output smem column
    int nvalid;  // outputs actually written (1..4)
    int flags;   // CG_*
    int pad0, pad1;
};
enum { CG_RELU = 1, CG_MASK = 2, CG_ACCUM = 4 };

// one lower-half column of one node's coupling (hint.py:79-84)
struct Ep {
    int x_col;   // column inside the x tile (0..d-1)
    int s_col;   // absolute smem column of this output's s value
    int t_col;   // absolute smem column of this output's t value
    int pad;
};

// weight-gradient job: dW[N x (K+1)] += dOut^T [N x TM] * [In | 1] [TM x (K+1)]   (bias = last column)
// Work item = (n-group of 4 rows, k-block of 16 columns, k-lane 0..3); a thread owns rows 4*ng..4*ng+3
// and columns 16*kb + lane + 4*j (j=0..3), so the 4 lanes of a k-block read 4 consecutive shared-memory
// columns per step (bank-conflict free with the TM+4 column pitch).
struct DwJob {
    int item_begin;  // prefix sum of items (nN*nKB*4) over the jobs of one phase
    int nN;          // n-groups (of 4 rows)
    int nKB;         // k-blocks (of 16 columns)
    int n_col;       // smem column of dOut row 0 (rows are 4-padded)
    int in0, k0, in1, k1;
    int out_off;     // offset of dW [4*nN][ld] in the per-CTA partial-gradient buffer
    int ld;          // row pitch there (= 16*nKB)
    int b_off;       // offset of db [4*nN] there
    int bitem_begin; // prefix sum of bias items (4*nN) over the jobs of one phase
};

struct Stage {
    int cg_begin[8];  // phases: 0 L1, 1 L2, 2 L3, 3 G3 (dH2), 4 G2 (dH1), 5 G1 (dx_upper), 6 GC (dc); [7] = end
    int ep_begin, ep_end;
    int dw_begin[4];  // job ranges of dW3, dW2, dW1; [3] = end
    int dw_items[3];
    int dw_bitems[3];
};

struct Schedule {
    int TM = 0;          // samples per tile
    int ncols = 0;       // shared-memory columns (each TM+4 floats)
    int DX = 0;          // columns of the x tile (d + dc rounded up to 4)
    int col_x = 0, col_d = -1, col_one = -1, col_zero = -1, col_out = 0, col_h1 = 0, col_h2 = 0;
    int raw_off = 0;     // float offset of the per-sample scratch (log-det partials / dJ) after the columns
    size_t smem_bytes = 0;
    std::vector<Stage> stages;   // root level first (inverse / backward order); forward walks it reversed
    std::vector<CG> cgs;
    std::vector<Ep> eps;
    std::vector<DwJob> dwjobs;
};

struct NodePack {      // packed-buffer offsets of one node
    int wt[2][3];      // W_l^T  [K_l][Np_l]   (forward operand)
    int b[2][3];       // bias   [Np_l]
    int wc3[2];        // W3 canonical [cout][hp]  (dH2 = dOut * W3)
    int wc2[2];        // W2 canonical [h][hp]
    int wg1;           // [2*hp][kp]: rows = s hidden units then t hidden units, cols = x_upper inputs
    int dw[2][3];      // offsets in the partial-gradient buffer: dW [Np_l][round16(K_l)]
    int db[2][3];      // ... and db [Np_l]
};

struct Plan {
    int d = 0, dc = 0;
    double clamp = 4.0;
    float alpha = 0.f;                 // (float)(clamp * 0.636)   hint.py:57,60
    int max_splits = -1, min_split_size = 2;
    std::vector<int> widths;
    std::vector<hint_node_info_t> nodes;       // pre-order
    std::vector<int64_t> param_offsets;        // [node][net][layer][kind] -> canonical flat offset
    int64_t n_params = 0;
    int64_t flops = 0;
    int max_depth = 0;

    std::vector<NodePack> packs;
    int64_t n_packed = 0;                      // floats in the packed weight buffer
    std::vector<int32_t> pack_src;             // packed[i] = pack_src[i] < 0 ? 0 : params[pack_src[i]]
    int64_t n_partial = 0;                     // floats in one CTA's partial-gradient buffer
    std::vector<int32_t> unpack_src;           // dparams[i] = sum_cta partial[cta][unpack_src[i]]

    Schedule fwd, bwd;
};

// Builds everything above.  Returns an empty string on success, else an error message; `code`
// receives HINT_ERR_*.
std::string build_plan(Plan& p, int d, int dc, const int32_t* c_internal, int n_internal, double clamp,
                       int max_splits, int min_split_size, int reshuffle, int* code);

inline int round4(int v) { return (v + 3) & ~3; }
inline int round8(int v) { return (v + 7) & ~7; }
inline int round16(int v) { return (v + 15) & ~15; }

}  // namespace hint
