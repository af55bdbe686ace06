// Plan of the register-chained warp-MMA kernels ("chain" kernels, chain_kernels.cuh) for one HINT block.
//
// Why a third tensor-core path: the interpreter kernels of plan_mma.h spend 97 % of their issue slots on op decoding,
// operand-ring moves and shared-memory round trips of the hidden activations (profiles/ncu_r01_bwd_mma_summary.txt:
// HMMA = 3.1 % of the executed instructions).  The subnets of hint.py:10-13 are three dependent dense layers whose hidden
// widths are small (h <= 72 for the d=42/43 `hint_8` and lens models), so one warp can carry a tile of samples through
// the WHOLE subnet in registers: with the k-slot permutation (slot t <-> feature 2t, slot t+4 <-> feature 2t+1) the C
// fragment of mma.sync.m16n8k8 IS the A fragment of the next layer (a = {c0, c2, c1, c3}), so bias, ReLU and the tf32
// rounding happen in place and no hidden activation ever touches shared memory.  Every shape is a compile-time
// constant (a node runs on the smallest instantiated shape that contains it, its operands zero-padded to that shape), so
// all loops unroll and operand addresses are immediates.
//
// Super nodes: most nodes of a HINT tree are tiny (24 of the 31 nodes of the d=43 tree have h = 8, 1-3 inputs and 1-3
// outputs) and would run one almost empty MMA per layer each.  Up to four adjacent tiny nodes of one tree level (they are
// independent) are therefore fused into one "super node": their inputs share the single k-step of layer 1, their hidden
// units are the n-tiles of one block-diagonal layer 2 (only the diagonal fragments are stored and multiplied), and their
// outputs share ONE n-tile of layer 3 - so the coupling epilogue, the operand loads and every per-node overhead are paid
// once per group, with the same number of MMAs.  A node is described by two small column maps (input feature -> tile
// column, output column -> coupled x column), which also cover the ordinary single-node case.
//
// Forward / inverse (hint.py:62-101): warps are independent - each owns a tile of 16*MT samples (x columns in a private
// shared-memory tile [column][sample]) for the whole tree, no CTA barrier anywhere.  Operands are staged in shared memory
// when they fit next to the tiles, else read through L1.  Backward: see chain_kernels.cuh.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

#if defined(__CUDACC__)
#define HINT_HD __host__ __device__
#else
#define HINT_HD
#endif

namespace hint {

constexpr int kChainMaxNodes = 128;
constexpr int kChainMaxIn = 24, kChainMaxOut = 24;    // 8 * max ks1, 8 * max no

// instantiated node shapes: k-steps of layer 1 (8 input features each), n-tiles of the hidden layers, n-tiles of the
// output, bd = layer 2 is block diagonal with 8x8 blocks (super node of `nh` tiny nodes)
struct ChainShape { int ks1, nh, no, bd; };
constexpr int kChainNumShapes = 9;
constexpr ChainShape kChainShapes[kChainNumShapes] = {{1, 1, 1, 0}, {1, 2, 1, 1}, {1, 4, 1, 1}, {1, 2, 1, 0}, {1, 3, 1, 0},
                                                      {1, 5, 1, 0}, {2, 5, 2, 0}, {2, 9, 2, 0}, {3, 9, 3, 0}};

// Packed operands.  Forward region (all nodes back to back; small enough to be staged in shared memory by the forward
// kernel): per node two nets (s, t), each [W1 | b1 | W2 | b2 | W3 | b3].  Transposed region (dgrad GEMMs of the backward),
// after the forward region: per node two nets, each [W3T | W2T | W1T].
// W* are B-fragment ordered (k-step major, n-tile, lane, 2 floats; a block-diagonal W2 keeps its nh diagonal fragments);
// b* natural order (a lane fetches (b[2t], b[2t+1]) of an n-tile with one 64-bit load: the C operand of the first k-step).
template <int KS1, int NH, int NO, int BD>
struct ChainOff {
    static constexpr int w2f = (BD ? NH : NH * NH) * 64;
    static constexpr int w1 = 0, b1 = KS1 * NH * 64, w2 = b1 + NH * 8, b2 = w2 + w2f, w3 = b2 + NH * 8, b3 = w3 + NH * NO * 64;
    static constexpr int net = b3 + NO * 8;                       // floats of one net in the forward region
    static constexpr int w3t = 0, w2t = NO * NH * 64, w1t = w2t + w2f;
    static constexpr int tnet = w1t + NH * KS1 * 64;              // ... in the transposed region
    // partial gradients of one net: [dW1T | dW2T | dW3T], dW_lT = [In | 1]^T dOut stored as C fragments (m-tile over IN
    // features with the bias row at index 8*k_tiles right after the padded features, n-tile over OUT features)
    static constexpr int mt1 = (8 * KS1 + 1 + 15) / 16, mth = (8 * NH + 1 + 15) / 16;
    static constexpr int dw1 = 0, dw2 = mt1 * NH * 128, dw3 = dw2 + mth * NH * 128;
    static constexpr int dnet = dw3 + mth * NO * 128;
};
struct ChainOffRt { int w1, b1, w2, b2, w3, b3, net, w3t, w2t, w1t, tnet, dw1, dw2, dw3, dnet; };
inline ChainOffRt chain_off(const ChainShape& s) {
    ChainOffRt o{};
    const int w2f = (s.bd ? s.nh : s.nh * s.nh) * 64;
    o.w1 = 0; o.b1 = s.ks1 * s.nh * 64; o.w2 = o.b1 + s.nh * 8; o.b2 = o.w2 + w2f; o.w3 = o.b2 + s.nh * 8;
    o.b3 = o.w3 + s.nh * s.no * 64; o.net = o.b3 + s.no * 8;
    o.w3t = 0; o.w2t = s.no * s.nh * 64; o.w1t = o.w2t + w2f; o.tnet = o.w1t + s.nh * s.ks1 * 64;
    const int mt1 = (8 * s.ks1 + 1 + 15) / 16, mth = (8 * s.nh + 1 + 15) / 16;
    o.dw1 = 0; o.dw2 = mt1 * s.nh * 128; o.dw3 = o.dw2 + mth * s.nh * 128; o.dnet = o.dw3 + mth * s.no * 128;
    return o;
}

struct ChainNode {      // device-visible, 128 bytes
    int shape;          // index into kChainShapes
    int w_off;          // float offset of the node's forward operands
    int wt_off;         // float offset of the node's transposed operands
    int dw_off;         // float offset of the node's partial-gradient block
    short in_col[kChainMaxIn];     // tile column of input feature f: x column, d + j for condition column j, -1 none
    short out_col[kChainMaxOut];   // x column coupled by output column c (hint.py:79-84), -1 none
    int pad[4];
};
static_assert(sizeof(ChainNode) == 128, "node records are 128 bytes");

struct ChainParam {     // travels as a __grid_constant__ kernel parameter
    ChainNode nodes[kChainMaxNodes];   // forward order: deepest tree level first, so children precede their parent
                                       // (hint.py:70-73); inverse and backward walk it reversed (hint.py:85-88)
};

struct ChainPlan {
    bool ok = false;
    std::string why;
    int n_nodes = 0;                    // (super) nodes
    ChainParam param;
    int64_t n_packed = 0;               // floats of the whole packed operand buffer
    int64_t n_fwd_packed = 0;           // floats of its forward region (a multiple of 4)
    std::vector<int32_t> pack_src;      // packed[i]: s >= 0 weight params[s] (rounded to tf32), -1 zero, s <= -2 bias params[-s-2] (exact)
    int64_t n_partial = 0;
    std::vector<int32_t> unpack_src;    // dparams[i] = sum_cta partial[cta][unpack_src[i]]
    int max_nh = 0, max_no = 0;          // n-tile slots of the backward kernel's hidden / output-gradient buffers
};

void build_chain_plan(const Plan& p, ChainPlan& c);

}  // namespace hint
