// Schedule of the warp-MMA (mma.sync.m16n8k8 tf32) fused-tree kernels for one HINT block.
//
// Why a second tensor-core path next to the tcgen05 one (plan_tc3.h): the coupling tree of hint.py:25-54 is made of
// many SMALL dense layers (h = 67/33/16/8/8 for the d=43 model).  tcgen05.mma is issue-bound for N <= 64
// (measured 38 cycles per MMA whatever N, profiles/ubench3_r01_mma_issue_tmem.txt) and its TMEM-resident tile
// serialises the 15 dependent layer->epilogue round trips of one tile (profiles/tc2_cycle_breakdown_r01.txt),
// while mma.sync sustains 1019 flop/cycle/SM (profiles/ubench5_r01_mma_sync_tf32.txt) with a 16x8x8 granularity
// that wastes little on the ragged widths and lets the backward's three GEMM shapes share one operand layout.
//
// Structure = the FP32 kernels' (plan.h): one CTA owns a tile of TM samples for the whole tree, the tile state is
// shared-memory columns [column][sample] (pitch TM+4), a stage = the nodes of one level that fit together, a phase
// = one layer of all those nodes followed by a CTA barrier.  What changes is the GEMM engine: every phase is a
// static list of WARP TASKS, each a register-tiled block of m16n8k8 MMAs, balanced over the CTA's warps by the
// planner.  Widths are padded to 8 (one n-tile); the k index inside each group of 8 is permuted (slot t <-> feature
// 2t, slot t+4 <-> feature 2t+1) so that A-fragment loads AND C-fragment stores are bank-conflict free at pitch
// TM+4, and weights are pre-packed in B-fragment order (one 64-bit load per lane per MMA).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace hint {

constexpr int kMmaThreads = 256;
constexpr int kMmaWarps = kMmaThreads / 32;
constexpr int kNC = 3;    // max n-tiles (8 output columns each) of a forward-type task; its m extent is the whole tile
constexpr int kPF = 3;    // k-steps of B fragments a forward-type task keeps in flight (weights stream from L2)
constexpr int kDMT = 2;   // max m-tiles (16 out-features each) of a weight-gradient task
constexpr int kDNC = 4;   // max n-tiles (8 in-features each) of a weight-gradient task

enum { MT_RELU = 1, MT_MASK = 2, MT_ACCUM = 4, MT_FAST = 8, MT_SYNC = 16 };   // MT_SYNC: CTA barrier before this op

// A warp's PROGRAM for one tile is a linear stream of 32-byte op records that the kernel interprets with the next record
// always prefetched.  All shared-memory positions are ELEMENT offsets (column * (TM+4)) so the kernel does no multiplies.
// A phase boundary is the MT_SYNC flag of the warp's next op (or a standalone OP_SYNC when it has none in that phase): the
// interpreter issues the next task's first weight fragments BEFORE that barrier, so L2 latency overlaps the wait.
enum { OP_END = 0, OP_GEMM = 1, OP_DW = 2, OP_SYNC = 3, OP_COUPLE = 4 };

// Forward-type task:  out[all TM rows] x [8*nt cols]  (op)=  in[rows x K] * B  (+ bias)
//   B fragment of (k-step ks, n-tile j of the chunk) = 64 floats at w_off + ks*ks_stride + j*64, lane l owns [2l, 2l+1]
//   OP_COUPLE reuses the record with w_off = first, b_off = one-past-last entry of eps[]
struct MTask {
    int w_off;
    int b_off;                            // bias of the chunk's first column (natural order) or -1
    unsigned short ks_stride, ksteps;
    unsigned short in_off, k0;            // input = columns at in_off (k0 of them) ++ columns at in1_off (k1); beyond: zero column
    unsigned short in1_off, k1;
    unsigned short out_off, zero_off;
    unsigned short nvalid;                // output columns actually stored (<= 8*nt)
    unsigned char mt, nt;
    unsigned char flags, pad0, pad1, type;
};
static_assert(sizeof(MTask) == 32, "op records are 32 bytes");

// Weight-gradient task:  dW[rows n0..) x [in-features kf0..)  (+)=  dOut^T * [In | 1]   contracted over the tile's samples
struct DTask {
    int out_off;                          // partial buffer: row r at out_off + r*ld, column = in-feature index (bias at column k0+k1)
    unsigned short a_off, N;              // dOut row 0; valid rows (rows >= N read the zero column and are not stored)
    unsigned short n0, kf0;
    unsigned short in_off, k0;            // in-feature f: f<k0 -> in_off+f cols; f<k0+k1 -> in1_off..; f==k0+k1 -> ones column
    unsigned short in1_off, k1;
    unsigned short ld, nstore;            // in-features stored: columns < nstore
    unsigned short one_off, zero_off;
    unsigned char flags, mt, nt, type;    // flags: MT_SYNC
};
static_assert(sizeof(DTask) == 32, "op records are 32 bytes");

struct WOp { int w[8]; };
static_assert(sizeof(WOp) == 32, "op records are 32 bytes");

// The op streams (and the coupling table) of ONE launch travel as a __grid_constant__ kernel parameter when they fit the
// 32 KB parameter space: the interpreter then fetches records through the constant cache instead of L1/L2 (with two
// 113 KB CTAs per SM only ~20 KB of L1 are left, and a missed record costs an L2 round trip on the critical path of every
// phase).  Larger programs stay in global memory (same interpreter, warp-uniform branch).
constexpr int kMaxProgOps = 928;
constexpr int kMaxEps = 256;
struct MmaParamProg {
    WOp ops[kMaxProgOps];
    unsigned short eps[kMaxEps][4];   // x_col, s_col, t_col, pad
};
static_assert(sizeof(MmaParamProg) <= 32000, "must leave room for the other kernel arguments in the 32764-byte parameter space");
enum { PROG_FWD = 0, PROG_INV = 1, PROG_BWD = 0 };

enum {  // phases of a stage; forward kernels run PH_L1..PH_L3, backward all
    PH_L1 = 0, PH_L2, PH_L3, PH_DW3, PH_G3, PH_DW2, PH_G2, PH_DW1G1, PH_COUNT
};

struct MStage {
    // tasks of phase p for warp w: [task_begin[p][w], task_begin[p][w+1]); dW phases index dtasks, the others mtasks;
    // PH_DW1G1 has both kinds (g_begin for the MTasks of G1/GC)
    int task_begin[PH_COUNT][kMmaWarps + 1];
    int g_begin[kMmaWarps + 1];
    int ep_begin, ep_end;
};

struct MSchedule {
    bool ok = false;
    std::string why;
    int TM = 0, ncols = 0;
    int col_x = 0, col_d = -1, col_one = -1, col_zero = -1, col_out = 0, col_h1 = 0, col_h2 = 0;
    int raw_off = 0;            // per-thread log-det partials (fwd) / per-sample dJ (bwd)
    size_t smem_bytes = 0;
    int ctas_per_sm = 1;
    std::vector<MStage> stages; // root level first (planner-internal: linearised into `prog`)
    std::vector<MTask> mtasks;
    std::vector<DTask> dtasks;
    std::vector<Ep> eps;
    std::vector<WOp> prog;      // all warps' op streams back to back
    int prog_begin[2][kMmaWarps] = {};   // [PROG_FWD | PROG_INV (forward schedule), PROG_BWD (backward schedule)][warp]
    int prog_end[2] = {};       // one past the last record of each program
    bool fits_param[2] = {};    // the program [prog_begin[p][0], prog_end[p]) and eps fit MmaParamProg
};

struct MmaPlan {
    bool ok = false;
    std::string why;
    int64_t n_packed = 0;               // floats of the packed operand buffer (B fragments + biases); the 3xTF32 mode keeps a
                                        // second buffer of the same layout with the low parts
    std::vector<int32_t> pack_src;      // packed[i] = pack_src[i] < 0 ? 0 : params[pack_src[i]]
    int64_t n_partial = 0;              // floats of one CTA's partial-gradient buffer
    std::vector<int32_t> unpack_src;    // dparams[i] = sum_cta partial[cta][unpack_src[i]]
    MSchedule fwd, bwd;
};

void build_mma_plan(const Plan& p, MmaPlan& m);

}  // namespace hint
