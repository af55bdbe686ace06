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
// Forward / inverse (hint.py:62-101): warps are independent - each owns a tile of 16*MT samples (x columns in a private
// shared-memory tile [column][sample]) for the whole tree, no CTA barrier anywhere.  Weights are read as B fragments
// straight from the packed operand buffer through L1 (the buffer is smaller than the L1 that is left).
// Backward: see the section in chain_kernels.cuh.
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

constexpr int kChainWarps = 8;
constexpr int kChainThreads = 32 * kChainWarps;
constexpr int kChainMaxNodes = 192;

// instantiated node shapes: k-steps of layer 1 (8 input features each), n-tiles of the hidden layers, n-tiles of the output
struct ChainShape { int ks1, nh, no; };
constexpr int kChainNumShapes = 7;
constexpr ChainShape kChainShapes[kChainNumShapes] = {{1, 1, 1}, {1, 2, 1}, {1, 3, 1}, {1, 5, 1}, {2, 5, 2}, {2, 9, 2}, {3, 9, 3}};

// Packed operands.  Forward region (all nodes back to back; small enough to be staged in shared memory by the forward
// kernel): per node two nets (s, t), each [W1 | b1 | W2 | b2 | W3 | b3].  Transposed region (dgrad GEMMs of the backward),
// after the forward region: per node two nets, each [W3T | W2T | W1T].
// W* are B-fragment ordered (k-step major, n-tile, lane, 2 floats), b* natural order.
HINT_HD constexpr int chain_w1(int, int, int) { return 0; }
HINT_HD constexpr int chain_b1(int ks1, int nh, int) { return ks1 * nh * 64; }
HINT_HD constexpr int chain_w2(int ks1, int nh, int no) { return chain_b1(ks1, nh, no) + nh * 8; }
HINT_HD constexpr int chain_b2(int ks1, int nh, int no) { return chain_w2(ks1, nh, no) + nh * nh * 64; }
HINT_HD constexpr int chain_w3(int ks1, int nh, int no) { return chain_b2(ks1, nh, no) + nh * 8; }
HINT_HD constexpr int chain_b3(int ks1, int nh, int no) { return chain_w3(ks1, nh, no) + nh * no * 64; }
HINT_HD constexpr int chain_net_floats(int ks1, int nh, int no) { return chain_b3(ks1, nh, no) + no * 8; }
HINT_HD constexpr int chain_w3t(int, int, int) { return 0; }
HINT_HD constexpr int chain_w2t(int, int nh, int no) { return no * nh * 64; }
HINT_HD constexpr int chain_w1t(int ks1, int nh, int no) { return chain_w2t(ks1, nh, no) + nh * nh * 64; }
HINT_HD constexpr int chain_tnet_floats(int ks1, int nh, int no) { return chain_w1t(ks1, nh, no) + nh * ks1 * 64; }

// Partial-gradient block of one node = two nets, each [dW1T | dW2T | dW3T] where dW_lT = [In | 1]^T * dOut is stored as the
// C fragments of that product: (m-tile over IN features, the bias row at index 8*k_tiles8 right after the padded features;
// n-tile over OUT features; lane; 4 floats).
HINT_HD constexpr int chain_dw_mt(int k_tiles8) { return (8 * k_tiles8 + 1 + 15) / 16; }
HINT_HD constexpr int chain_dw1(int, int, int) { return 0; }
HINT_HD constexpr int chain_dw2(int ks1, int nh, int) { return chain_dw_mt(ks1) * nh * 128; }
HINT_HD constexpr int chain_dw3(int ks1, int nh, int no) { return chain_dw2(ks1, nh, no) + chain_dw_mt(nh) * nh * 128; }
HINT_HD constexpr int chain_dw_net_floats(int ks1, int nh, int no) { return chain_dw3(ks1, nh, no) + chain_dw_mt(nh) * no * 128; }

struct ChainNode {      // device-visible, 32 bytes
    int shape;          // index into kChainShapes
    int lo, k, cout;    // upper half = columns [lo, lo+k), lower half = [lo+k, lo+k+cout)   (hint.py:41,68)
    int cin;            // k + dc (hint.py:44)
    int w_off;          // float offset of the node's forward operands
    int wt_off;         // float offset of the node's transposed operands
    int dw_off;         // float offset of the node's partial-gradient block
};

struct ChainParam {     // travels as a __grid_constant__ kernel parameter
    ChainNode nodes[kChainMaxNodes];   // forward order: children before their parent (hint.py:70-73); inverse and
                                       // backward walk it reversed (hint.py:85-88)
};

struct ChainPlan {
    bool ok = false;
    std::string why;
    int n_nodes = 0;
    ChainParam param;
    int64_t n_packed = 0;               // floats of the whole packed operand buffer
    int64_t n_fwd_packed = 0;           // floats of its forward region (a multiple of 4)
    std::vector<int32_t> pack_src;      // packed[i]: s >= 0 weight params[s] (rounded to tf32), -1 zero, s <= -2 bias params[-s-2] (exact)
    int64_t n_partial = 0;
    std::vector<int32_t> unpack_src;    // dparams[i] = sum_cta partial[cta][unpack_src[i]]
    int max_nh = 0, max_no = 0;
};

void build_chain_plan(const Plan& p, ChainPlan& c);

}  // namespace hint
