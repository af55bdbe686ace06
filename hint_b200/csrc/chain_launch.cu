// Launch glue of the register-chained warp-MMA kernels.
#include "chain_launch.h"
#include "launch_count.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chain_kernels.cuh"

namespace hint {

namespace {

// forward / inverse configuration: MT = 1 (16 samples per warp), 12 (or 8) warps next to the shared-memory resident
// operands (WS), else 16 warps reading the operands through L1.  HINT_B200_CHAIN_FWD=<mt><nw><ws> (e.g. "2080",
// "1161") overrides (developer aid).
// developer switches exist only in -DHINT_B200_DEV builds (see HINT_EXP in chain_kernels.cuh)
const char* dev_getenv(const char* name) {
#ifdef HINT_B200_DEV
    return std::getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}
int chain_exp() {
    static const int v = [] { const char* e = dev_getenv("HINT_B200_CHAIN_EXP"); return e ? std::atoi(e) : 0; }();
    return v;
}

struct FwdCfg { int mt, nw, ws; size_t smem; };

template <int MT, int NW, bool WS>
cudaError_t set_attr() {
    cudaError_t e = cudaFuncSetAttribute((const void*)hint_fwd_chain_kernel<MT, NW, false, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute((const void*)hint_fwd_chain_kernel<MT, NW, true, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
}

size_t fwd_smem(const Plan& p, const ChainPlan& c, int mt, int nw, int ws) {
    return mt == 2 ? chain_fwd_smem_bytes<2>(c.n_nodes, p.d, p.dc, nw, c.n_fwd_packed, ws != 0)
                   : chain_fwd_smem_bytes<1>(c.n_nodes, p.d, p.dc, nw, c.n_fwd_packed, ws != 0);
}

FwdCfg pick_fwd(const Plan& p, const ChainPlan& c) {
    static const char* e = dev_getenv("HINT_B200_CHAIN_FWD");
    if (e && e[0] && e[1] && e[2] && e[3]) {
        FwdCfg f{e[0] - '0', 10 * (e[1] - '0') + (e[2] - '0'), e[3] - '0', 0};
        f.smem = fwd_smem(p, c, f.mt, f.nw, f.ws);
        return f;
    }
    // 32 samples per warp (MT = 2: half the B-fragment loads per sample, 218 registers, 8 warps) was measured too: 1.21 ms forward /
    // 0.89 ms inverse against 0.95 / 0.89 ms for MT = 1 with 12 warps - fewer warps lose more than the operand reuse gains
    for (int nw : {12, 8}) {      // measured (d=43): 8 warps 1.52 ms, 12 warps 0.96 ms, 14 warps 1.00 ms - beyond 12 the shared-memory
                                  // pipe (one 256-byte B fragment per MMA at 16 samples per warp) is saturated
        const size_t b = fwd_smem(p, c, 1, nw, 1);
        if (b <= (size_t)kSmemMax) return FwdCfg{1, nw, 1, b};
    }
    return FwdCfg{1, 16, 0, fwd_smem(p, c, 1, 16, 0)};
}

// backward configuration: (MT, NW) = (1, 4), two 107 KB CTAs per SM.  Measured and dropped (kept in the CPU emulation tests as
// geometry variants): (1, 8) one CTA per SM 5.2 ms, (2, 4) 7.3 ms vs 5.0 ms.
struct BwdCfg { int mt, nw; };
BwdCfg pick_bwd() { return BwdCfg{1, 4}; }
ChainBwdSmem bwd_layout(const Plan& p, const ChainPlan& c, BwdCfg) {
    return chain_bwd_smem<1, 4>(p.d, p.dc, c.max_nh, c.max_no, c.n_nodes);
}
template <int MT, int NW, int MINB>
cudaError_t setup_bwd(size_t smem, int num_sms, int* ctas) {
    const void* fn = (const void*)hint_bwd_chain_kernel<MT, NW, MINB>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32 * NW, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    *ctas = occ * num_sms;
    return cudaSuccess;
}

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& v) {
    *dst = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

bool chain_fits(const Plan& p, const ChainPlan& c, std::string* why) {
    const size_t f = fwd_smem(p, c, 1, 16, 0), b = (size_t)bwd_layout(p, c, pick_bwd()).total * 4;
    if (f <= (size_t)kSmemMax && b <= (size_t)kSmemMax) return true;
    if (why) *why = "tile state of the register-chained kernels (" + std::to_string(std::max(f, b)) + " B) exceeds shared memory";
    return false;
}

cudaError_t chain_setup(const Plan& p, const ChainPlan& c, int num_sms, DevChain& d) {
    cudaError_t e;
    d.num_sms = num_sms;
    const FwdCfg f = pick_fwd(p, c);
    d.fwd_smem = f.smem;
    if (d.fwd_smem > (size_t)kSmemMax) return cudaErrorInvalidValue;
    if ((e = set_attr<1, 12, true>()) != cudaSuccess) return e;
    if ((e = set_attr<1, 8, true>()) != cudaSuccess) return e;
    if ((e = set_attr<1, 16, false>()) != cudaSuccess) return e;
    const BwdCfg b = pick_bwd();
    d.bwd_mt = b.mt; d.bwd_nw = b.nw;
    d.bwd_smem = (size_t)bwd_layout(p, c, b).total * 4;
    if (d.bwd_smem > (size_t)kSmemMax) return cudaErrorInvalidValue;
    if ((e = setup_bwd<1, 4, 2>(d.bwd_smem, num_sms, &d.bwd_ctas)) != cudaSuccess) return e;
    if ((e = upload(&d.pack_src, c.pack_src)) != cudaSuccess) return e;
    return upload(&d.unpack_src, c.unpack_src);
}

void chain_free(DevChain& d) {
    cudaFree(d.pack_src); cudaFree(d.unpack_src);
    d.pack_src = d.unpack_src = nullptr;
}

cudaError_t chain_pack(const ChainPlan& c, const DevChain& d, const float* params, float* packed, cudaStream_t st) {
    const long long n = c.n_packed;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
    hint_pack_mma_kernel<<<blocks, threads, 0, st>>>(d.pack_src, params, packed, nullptr, n, 1, 0); HINT_LAUNCHED();
    return cudaGetLastError();
}

int chain_bwd_ctas(const ChainPlan& c, const DevChain& d, long long B) {
    (void)c;
    const int TM = 16 * d.bwd_mt * d.bwd_nw;
    const long long ntiles = (B + TM - 1) / TM;
    return (int)std::min<long long>(ntiles, d.bwd_ctas);
}

cudaError_t chain_launch_bwd(const Plan& p, const ChainPlan& c, const DevChain& d, int grid, const float* z, const float* cond,
                             const float* packed, const float* dz, const float* dlogdet, float* x_rec, float* dx, float* dc,
                             float* partials, long long B, cudaStream_t st, float nll_scale) {
    ChainTables T{c.n_nodes, p.d, p.dc, p.alpha, (int)c.n_fwd_packed, chain_exp(), nll_scale};
    const BwdCfg b{d.bwd_mt, d.bwd_nw};
    const ChainBwdSmem L = bwd_layout(p, c, b);
    const long long np = c.n_partial;
    hint_bwd_chain_kernel<1, 4, 2><<<grid, 128, d.bwd_smem, st>>>(T, c.param, L, z, cond, packed, dz, dlogdet, x_rec, dx, dc, partials, np, B); HINT_LAUNCHED();
#ifdef HINT_B200_DEV
    if (T.exp & 32) {   // developer aid: phase-boundary cycle stamps of CTA 0 (HINT_B200_CHAIN_EXP=32)
        static long long h[2048];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_chain_dbg, sizeof(h));
        std::fprintf(stderr, "[hint_b200 chain dbg] %lld stamps; deltas:", h[0]);
        for (long long i = 1; i < h[0]; ++i) std::fprintf(stderr, " %lld", h[1 + i] - h[i]);
        std::fprintf(stderr, "\n");
        std::memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_chain_dbg, h, sizeof(h));
    }
#endif
    return cudaGetLastError();
}

cudaError_t chain_launch_fwd(const Plan& p, const ChainPlan& c, const DevChain& d, const float* x, const float* cond,
                             const float* packed, float* z, float* logdet, long long B, int rev, cudaStream_t st) {
    ChainTables T{c.n_nodes, p.d, p.dc, p.alpha, (int)c.n_fwd_packed, chain_exp()};
    const FwdCfg f = pick_fwd(p, c);
    const int RW = 16 * f.mt;
    const long long ntiles = (B + RW - 1) / RW;
    const int grid = (int)std::min<long long>((ntiles + f.nw - 1) / f.nw, d.num_sms);
#define HINT_CHAIN_LAUNCH(MT, NW, WS)                                                                                          \
    if (f.mt == MT && f.nw == NW && f.ws == WS) {                                                                              \
        if (rev) { hint_fwd_chain_kernel<MT, NW, true, WS != 0><<<grid, 32 * NW, f.smem, st>>>(T, c.param, x, cond, packed, z, logdet, B); HINT_LAUNCHED(); }  \
        else { hint_fwd_chain_kernel<MT, NW, false, WS != 0><<<grid, 32 * NW, f.smem, st>>>(T, c.param, x, cond, packed, z, logdet, B); HINT_LAUNCHED(); }     \
        return cudaGetLastError();                                                                                             \
    }
    HINT_CHAIN_LAUNCH(1, 12, 1) HINT_CHAIN_LAUNCH(1, 8, 1) HINT_CHAIN_LAUNCH(1, 16, 0)
#undef HINT_CHAIN_LAUNCH
    return cudaErrorInvalidValue;

}

}  // namespace hint
