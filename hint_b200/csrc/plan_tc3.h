// Planner of the tcgen05 / TMEM TRAINING kernel ("tc3": memory-free backward of one HINT block, hint.py:62-101 +
// the autograd tape over it, in TF32 on the 5th-generation tensor cores).  Pure C++ (no CUDA): unit-tested on CPU
// through tests/emul/emul_tc3.cpp, which interprets the same tables.
//
// Machine model (one CTA = one tile of 128 samples, persistent over tiles; TMEM base column 0, all 512 columns):
//   * chain GEMMs run with M = 128 SAMPLES on the TMEM lanes: A = activations in TMEM (tcgen05.mma, A from TMEM),
//     B = tf32 weights streamed from L2 through a shared-memory ring (canonical un-swizzled K-major blocks),
//     D = TMEM columns.  Biases are one extra K step against a constant "ones" column block.
//   * weight-gradient GEMMs contract over the tile's 128 samples: both operands are K-major SWIZZLE_128B "images"
//     [feature][sample] that the epilogue threads (thread = sample) write conflict-free next to their TMEM results;
//     D = TMEM accumulators with M = 128 FEATURES on the lanes.  Accumulators are added to the CTA's private
//     partial-gradient buffer after every tile (coalesced, through a shared-memory staging transpose).
//   * nodes of one tree level are processed together as a GROUP: block-diagonal hidden layers, dense first/last
//     layers (zero blocks in the packed weights), one weight-gradient GEMM per layer and group.
//   * per group three phases: S forward (keeps only s), T forward + coupling backward + T backward, S forward again
//     (recompute) + S backward.  TMEM holds one net at a time.
//   * two roles walk static programs: the MMA issuer (records = K loops of tcgen05.mma) and 8 epilogue warps (steps).
//     Cross-role ordering is inferred by the planner from the read / write sets of the steps (wait indices), never
//     hand-written; tests run the program under an eager and a lazy MMA-completion emulator.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace hint {

constexpr int kT3EpiWarps = 16;                       // 4 warps per TMEM lane quadrant: 4 column slices of every wide step
constexpr int kT3Threads = 32 * (kT3EpiWarps + 2);   // + issuer warp + loader warp
constexpr int kT3NB = 16;                             // mbarriers per signal sequence
constexpr int kT3MaxSlots = 4;
constexpr int kT3Imgs = 5;                            // image buffers: 0..2 hidden images, 3 = subnet input, 4 = dOut

// ---- issuer program --------------------------------------------------------------------------------------------
enum : uint16_t {
    T3M_SS = 1,         // both operands from shared-memory images (weight gradient); else A from TMEM, B from the ring
    T3M_ZERO = 2,       // first K step overwrites D
    T3M_NEWCHUNK = 4,   // wait for the next ring slot before issuing
    T3M_ENDCHUNK = 8,   // tcgen05.commit -> the slot's "empty" barrier after issuing
    T3M_COMMIT = 16,    // tcgen05.commit -> next MMA signal after issuing
};
struct T3Mma {
    uint32_t idesc;      // kind::tf32 instruction descriptor (M = 128, N)
    uint32_t b_off;      // TS: byte offset of the B block inside the ring slot.  SS: a_img | b_img << 8
    uint16_t d_col;      // TMEM column of D
    uint16_t a_col;      // TS: TMEM column of A.  SS: first 8-row group of the A tile inside its image
    uint16_t nk;         // K steps of 8
    uint16_t b_sbo16;    // TS: stride between 8-row groups of B, in 16-byte units
    uint16_t flags;
    int16_t wait_epi;    // epilogue step (index inside the tile program) that must have completed, -1 = none
};

// ---- epilogue program ------------------------------------------------------------------------------------------
enum : int32_t {
    T3E_IN = 0,      // subnet inputs of a group: state columns -> TMEM A columns + input image
    T3E_HID,         // hidden layer: relu + tf32 rounding in place (+ image, + ones block)
    T3E_OUTS,        // phase 1: s outputs -> state
    T3E_CPL,         // phase 2: coupling inverse + its backward (uses saved s, fresh t); emits dOut of t, saves dOut of s
    T3E_DS,          // phase 3: saved dOut of s -> TMEM + image
    T3E_DHID,        // dgrad hidden layer: relu mask + rounding in place + image
    T3E_DA,          // input gradient: accumulate into the gradient state
    T3E_FLUSH,       // weight-gradient accumulator -> partial buffer
    T3E_CPLF,        // forward / inverse kernels: the coupling itself (uses saved s, fresh t) + log-det accumulation
};
// kinds of program build_tc3_plan() generates
enum : int { T3K_BACKWARD = 0, T3K_FORWARD = 1, T3K_INVERSE = 2 };
// all fields 32-bit: the kernel reads a step as three 128-bit loads and never unpacks sub-word fields (see the note at
// T3MmaWords in tc3_kernels.cuh)
struct T3Epi {
    int32_t type;
    int32_t flags;
    int32_t wait_mma;    // MMA signal (index inside the tile program) that must have fired, -1 = none
    int32_t off;         // partial-buffer offset (T3E_FLUSH, T3E_CPL, T3E_DS)
    int32_t a, b, c, d, e, f, g, h;   // per type, see plan_tc3.cpp: the emitters in build_tc3_plan()
};
// T3E_IN flags
enum : int32_t { T3I_NOIMG = 1 };   // transport programs: no input image
// T3E_HID flags
enum : int32_t { T3H_IMG = 1, T3H_ONES = 2, T3H_IMG_ONES = 4 };
// T3E_DHID flags
enum : int32_t { T3D_MASK_TMEM = 1 };   // relu mask from TMEM (else from an image)
// T3E_FLUSH kinds (field g): which column range of a lane's node is flushed
enum : int32_t { T3F_W2 = 0, T3F_W1 = 1, T3F_W3 = 2 };

// Biases enter the chain GEMMs as one extra K step against a constant block whose first TWO columns are 1: the packed operand
// holds tf32(b) in the first and tf32(b - tf32(b)) in the second (pack_src entry | kT3BiasLo), so a bias keeps ~21 mantissa
// bits instead of 10 - a rounded output-layer bias would shift t (and z) by up to 2^-11 |b| on its own.
constexpr int32_t kT3BiasLo = 1 << 30;

struct T3Chunk {
    uint32_t g_off;      // float offset in the packed weight buffer
    uint32_t bytes;
};

// one group = same-depth nodes processed together
struct T3Group {
    std::vector<int> nodes;
    std::vector<int> hoff, xoff, ooff;   // per node: first hidden feature / input column / output column
    int HP = 0;      // hidden columns (each node padded to 16)
    int KX = 0;      // sum of k (x_upper columns); then dc condition columns; then the ones column
    int KA = 0;      // padded to 8
    int OC = 0, KD = 0, OW = 0;   // outputs: true, padded to 8 (K of dOut), padded to 16 (N of layer 3)
    int N2 = 0;      // N of the dW2 GEMM: pad16(HP + 8)
    int N1 = 0;      // N of the dW1 / dA GEMMs: pad16(KX + dc + 1)
    int mtiles = 1;  // M tiles of the weight-gradient GEMMs
    int CH = 0;      // transport programs, wide single node: the second hidden layer is produced and consumed in chunks of CH columns
    int part[2][4];  // partial-buffer offsets per net: dW2 block, dW1 block, dW3 block, db3 vector
    // TMEM map of this group (columns)
    int tm_p = 0, tm_q = 0, tm_acc2 = 0, tm_ain = 0, tm_dout = 0, tm_out = 0, tm_da = 0, tm_acc1 = 0, tm_acc3 = 0;
    int tab_nodes = 0;   // offset of the node table in tab16: per node {hoff, h, xoff, k, ooff, cout}
};

struct T3Plan {
    bool ok = false;
    std::string why;
    int kind = T3K_BACKWARD;
    int d = 0, dc = 0;
    float alpha = 0.f;
    std::vector<T3Group> groups;          // root level first
    // shared memory map (bytes from the start of dynamic shared memory)
    int sm_bars = 0, sm_tab16 = 0, sm_epis = 0, sm_xs = 0, sm_gs = 0, sm_os = 0, sm_stage = 0, sm_red = 0, sm_ring = 0;
    int sm_img[kT3Imgs] = {0, 0, 0, 0, 0};
    int img_rows[kT3Imgs] = {0, 0, 0, 0, 0};   // allocated rows (multiple of 8); slab = rows * 128 bytes per 32 samples
    int n_imgs_hidden = 2;                       // 2 or 3 hidden image buffers
    int xp = 0, op = 0;       // row pitch (floats) of the x / gradient state and of the s-output state (odd)
    int slot_bytes = 0, n_slots = 0;
    int smem_bytes = 0;
    // programs of one tile
    std::vector<T3Mma> mmas;
    std::vector<T3Epi> epis;
    std::vector<T3Chunk> chunks;
    std::vector<int16_t> tab16;   // small tables of the epilogue steps
    int n_mma_signals = 0;
    // weights
    int64_t n_packed = 0;                // floats
    std::vector<int32_t> pack_src;       // packed[i] = pack_src[i] < 0 ? 0 : tf32(params[pack_src[i]]); | kT3BiasLo: the tf32 residual
    int64_t n_partial = 0;               // floats per CTA
    std::vector<int32_t> unpack_src;     // dparams[i] = sum over CTAs of partial[unpack_src[i]]
    std::vector<uint8_t> unpack_q4;      // 1: the parameter (a layer-3 bias) is the sum of 4 per-quadrant slots at stride 32
    // model of one tile, for reports: MMA instructions and tensor-pipe cycles (N/2 per instruction)
    int64_t n_mma_instr = 0, tensor_cycles = 0;
};

// envelope: every group must fit TMEM (512 columns) and shared memory.  Never throws; on failure ok = false + why.
// kind = T3K_BACKWARD: the memory-free backward (training kernel).  T3K_FORWARD / T3K_INVERSE: the same machine running the
// transport alone (hint.py:62-101): per group S chain -> s, T chain -> t, coupling step; no images, no accumulators, deepest
// group first (forward, hint.py:70-73) or root first (inverse, hint.py:85-88).
void build_tc3_plan(const Plan& p, T3Plan& t, int kind = T3K_BACKWARD);

// SWIZZLE_128B K-major image: float index of (feature row r, sample s) in an image of `rows` allocated rows
inline int t3_img_off(int r, int s, int rows) {
    return (s >> 5) * rows * 32 + (r >> 3) * 256 + (r & 7) * 32 + ((((s & 31) >> 2) ^ (r & 7)) << 2) + (s & 3);
}
// canonical un-swizzled K-major block [N][K] (K multiple of 8): float index of (n, k)
inline int t3_canon_off(int n, int k, int K) { return (n >> 3) * (K * 8) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

}  // namespace hint
