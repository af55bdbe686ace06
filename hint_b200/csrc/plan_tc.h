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
    signed char wait_epi;     // job of THIS stage whose epilogue must be complete before issue, or -1
    signed char job;          // TC_J* this op belongs to
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

// ---- v2 program: the same schedule, re-encoded for the kernel of tc2_kernels.cuh ---------------------------------
// The whole program is passed BY VALUE as a __grid_constant__ kernel parameter, so the MMA-issuing threads fetch
// their operands with uniform constant-bank loads (measured: 38 cycles per tcgen05.mma with warp-uniform
// operands against 77 when the compiler has to wrap the instruction in a divergence "waterfall" loop;
// profiles/ubench3_r01_mma_issue_tmem.txt).  Ops are pre-partitioned per issuing warp, so nobody walks ops it
// does not own.
constexpr int kT2MaxOps = 1000, kT2MaxSegs = 288, kT2MaxStages = 48, kT2MaxFins = 400, kT2MaxChunks = 96, kT2MaxXw = 192;
enum { T2_FIRST_IN_JOB = 1, T2_LAST_IN_JOB = 2, T2_FIRST_IN_CHUNK = 4, T2_LAST_IN_CHUNK = 8 };

struct T2Op {                 // 16 bytes
    uint32_t da;              // d_col | a_col << 16        (TMEM columns; the CTA owns all 512, base 0)
    uint32_t b16;             // byte offset of the B image inside its ring slot, >> 4
    uint32_t sbo_nk;          // (SBO >> 4) | nk << 16      (SBO = 8-row group pitch of the canonical K-major image)
    uint32_t idesc;           // tcgen05 instruction descriptor; bit 0 = accumulate onto D from the first K step
};
struct T2Seg {                // maximal run of ops of one job reading one weight chunk; 14 bytes
    uint16_t op_ofs[kTcIssuers + 1];   // ops of issuer q: [op_ofs[q], op_ofs[q+1])
    uint8_t job, flags;
    uint16_t pad;
};
struct T2Hidden { uint16_t col0, ncols, bias_off, pad; };
struct T2Stage {
    uint16_t seg_begin, seg_end, fin_begin, fin_end, chunk_begin, chunk_end;
    T2Hidden hid[3];
};
struct T2Fin { uint16_t s_col, t_col, x_col, bs_off, bt_off; };   // bias offsets relative to the first bias float
struct T2Chunk { uint32_t g_off16, bytes; };                      // offset in the packed buffer (16-byte units), size
struct T2Prog {
    int nstages, nfins, d, dc, xw, xc, xr;
    int slot_bytes, n_slots, n_bias, bias_base;
    int smem_tab, smem_bias, smem_in, smem_in_bytes, smem_out, smem_ring;   // byte offsets in dynamic shared memory
    float alpha;
    int round_acts;
    T2Stage stages[kT2MaxStages];
    T2Seg segs[kT2MaxSegs];
    T2Op ops[kT2MaxOps];
    T2Fin fins[kT2MaxFins];
    T2Chunk chunks[kT2MaxChunks];
    int16_t xlog[kT2MaxXw];
};
static_assert(sizeof(T2Prog) <= 32000, "the program must fit the 32 KB kernel-parameter space");

struct T2Host {
    bool ok = false;
    std::string why;
    size_t smem_bytes = 0;
    T2Prog prog;
};
void build_tc2_program(const Plan& p, const TcSchedule& t, T2Host& out);

}  // namespace hint
