// Static program of the tcgen05 (TF32) fused-tree kernel for one HINT block.
//
// One CTA owns a tile of 128 samples (TMEM lane = sample).  The x tile, every hidden activation and every
// s/t output of the nodes being processed live in TMEM columns; torch.split / torch.cat of hint.py:68,90 are
// column bookkeeping.  Each subnet layer of each node is a group of tcgen05.mma (M=128, kind::tf32) whose A
// operand is read straight from TMEM (the previous layer's accumulator after an in-place bias+ReLU epilogue,
// or the x tile itself) and whose B operand (the weights, pre-packed in the un-swizzled K-major canonical
// layout) is streamed from L2 into a shared-memory ring by bulk copies.
//
// The plan below is the complete, static instruction stream of that machine: per stage (a set of nodes of one
// tree level) the MMA ops, the weight chunks they consume, and the epilogue descriptors.  The kernel is an
// interpreter of these tables; tests/emul interprets the same tables on the CPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace hint {

enum { TC_ACCUM = 1, TC_FIRST_IN_CHUNK = 2, TC_LAST_IN_CHUNK = 4 };
constexpr int kTcIssuers = 4;
enum { TC_J1 = 0, TC_J2S = 1, TC_J2T = 2, TC_J3S = 3, TC_J3T = 4, TC_NJOBS = 5 };

struct TcOp {                 // D[128 x N] (+)= A[128 x 8*nk] * B[N x 8*nk]^T
    int d_col, a_col;         // TMEM columns
    int b_off;                // float offset of the B image inside its weight chunk
    int nk;                   // K steps of 8
    unsigned idesc;           // tcgen05 instruction descriptor (tf32, M=128, N)
    int n_rows;               // N (for the emulator)
    int flags;                // TC_*
    short wait_epi;           // job of THIS stage whose epilogue must be complete before issue, or -1
    signed char commit_job;   // job completed by this op (every issuer commits -> mma_done[job]), or -1
    unsigned char issuer;     // which of the kTcIssuers MMA warps issues this op (issue cost ~80 cycles/MMA per thread)
};

struct TcChunk {              // contiguous piece of the packed weight image, copied into one ring slot
    long long g_off;          // float offset in the TC packed buffer
    int bytes;
    int pad;
};

struct TcHidden {             // in-place epilogue v = relu(v + bias) on TMEM columns [col0, col0+ncols)
    int col0, ncols, bias_off, pad;
};

struct TcFinal {              // one 4-column group of one node's coupling (hint.py:79-84)
    int s_col, t_col, x_col;  // TMEM columns of s, t and of the lower-half x values
    int nvalid;               // 1..4 valid columns
    int bs_off, bt_off;       // float offsets of the s / t output biases (4 each) in the TC packed buffer
    int pad0, pad1;
};

struct TcStage {
    int op_begin, op_end;
    int chunk_begin, chunk_end;
    TcHidden hid[3];          // epilogues of jobs J1, J2S, J2T (ncols == 0: job absent)
    int has_job[TC_NJOBS];
    int fin_begin, fin_end;
    int pad;                  // keeps sizeof(TcStage) a multiple of 8 (tables are packed back to back in smem)
};
static_assert(sizeof(TcStage) % 8 == 0 && sizeof(TcOp) % 8 == 0 && sizeof(TcChunk) % 8 == 0 && sizeof(TcFinal) % 8 == 0,
              "table records must keep 8-byte alignment when concatenated");

struct TcSchedule {
    bool ok = false;
    std::string why;                    // reason when !ok (config outside the TF32 kernel's current envelope)
    int d = 0, dc = 0;
    int xw = 0;                         // physical width of the x columns in TMEM (leaves padded to 4)
    int xc = 0;                         // first TMEM column of the condition
    int xr = 0;                         // columns reserved for x + condition (multiple of 16)
    std::vector<int> xphys;             // logical column -> TMEM column
    std::vector<int> xlog;              // TMEM column (< xw) -> logical column or -1
    int tmem_cols = 512;
    int slot_bytes = 0, n_slots = 0;
    std::vector<TcStage> stages;        // root level first
    std::vector<TcOp> ops;
    std::vector<TcChunk> chunks;
    std::vector<TcFinal> fins;
    long long n_packed = 0;             // floats in the TC packed buffer (weight images + biases)
    std::vector<int32_t> pack_src;      // packed[i] = pack_src[i] < 0 ? 0 : params[pack_src[i]]
    long long n_weight_floats = 0;      // the first n_weight_floats entries are MMA operands (rounded to tf32)
    // shared-memory layout (byte offsets)
    int smem_stage_in = 0, smem_stage_bytes = 0;   // 2 x staging buffers of the x (+c) tile
    int smem_ring = 0;
    int smem_tables = 0, smem_tables_bytes = 0;
    int smem_bars = 0;
    size_t smem_bytes = 0;
};

// Builds the TF32 program for a plan (after build_plan).  Never fails hard: !ok + why when outside the envelope.
void build_tc_schedule(const Plan& p, TcSchedule& t);

inline int round8(int v) { return (v + 7) & ~7; }

}  // namespace hint
