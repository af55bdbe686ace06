// C ABI of hint_b200 (see include/hint_b200.h).  Host glue only: validates arguments, uploads the
// plan's descriptor tables once per device, and launches the kernels on the caller's stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "launch_count.h"
#include "plan.h"
#include "plan_mma.h"
#include "mma_launch.h"
#include "plan_chain.h"
#include "chain_launch.h"
#include "plan_tc3.h"
#include "tc3_launch.h"
#include "train_ops.h"
#include "householder.h"
#include "mlp_coupling.h"
#include "mmd.h"
#include "simt_kernels.cuh"

using namespace hint;

// Developer switches (kernel-family overrides, cycle-breakdown dumps) exist only in -DHINT_B200_DEV builds: in the product an
// environment variable cannot change which kernel runs.
static inline const char* dev_getenv(const char* name) {
#ifdef HINT_B200_DEV
    return std::getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

namespace hint {
std::atomic<unsigned long long> g_launches{0};
}

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(HINT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    } while (0)

struct DevSchedule {
    CG* cgs = nullptr;
    Ep* eps = nullptr;
    DwJob* dwjobs = nullptr;
    Stage* stages = nullptr;
    int max_ctas = 0;  // SMs x occupancy
};

struct DevPlan {
    DevSchedule fwd, bwd;
    DevMma mma;
    DevChain chain;
    DevTc3 tc3, tc3f, tc3i;   // training kernel; the same machine running the forward / inverse transport
    int* pack_src = nullptr;
    int* unpack_src = nullptr;
    int num_sms = 0;
};

}  // namespace

struct hint_plan {
    Plan p;
    MmaPlan mma;
    ChainPlan chain;
    T3Plan tc3, tc3f, tc3i;
    std::mutex mu;
    std::map<int, DevPlan> dev;  // per CUDA device ordinal
};

namespace {

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& v) {
    *dst = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

template <int TM>
cudaError_t setup_kernels(const Schedule& s, bool bwd, int num_sms, int* max_ctas) {
    const void* fn = bwd ? (const void*)hint_bwd_fp32_kernel<TM> : (const void*)hint_fwd_fp32_kernel<TM>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);  // per function, shared by all plans
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, s.smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    *max_ctas = occ * num_sms;
    return cudaSuccess;
}

cudaError_t setup_schedule(const Schedule& s, bool bwd, int num_sms, DevSchedule& ds) {
    cudaError_t e;
    if ((e = upload(&ds.cgs, s.cgs)) != cudaSuccess) return e;
    if ((e = upload(&ds.eps, s.eps)) != cudaSuccess) return e;
    if ((e = upload(&ds.dwjobs, s.dwjobs)) != cudaSuccess) return e;
    if ((e = upload(&ds.stages, s.stages)) != cudaSuccess) return e;
    switch (s.TM) {
        case 128: return setup_kernels<128>(s, bwd, num_sms, &ds.max_ctas);
        case 64: return setup_kernels<64>(s, bwd, num_sms, &ds.max_ctas);
        case 32: return setup_kernels<32>(s, bwd, num_sms, &ds.max_ctas);
        case 16: return setup_kernels<16>(s, bwd, num_sms, &ds.max_ctas);
        case 8: return setup_kernels<8>(s, bwd, num_sms, &ds.max_ctas);
    }
    return cudaErrorInvalidValue;
}

// Device-side tables for the current device (created on first use).
int get_dev(hint_plan* hp, DevPlan** out) {
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(hp->mu);
    auto it = hp->dev.find(dev);
    if (it != hp->dev.end()) { *out = &it->second; return HINT_OK; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(HINT_ERR_UNSUPPORTED, "hint_b200 kernels are built for sm_100a only; device is sm_" +
                                              std::to_string(prop.major) + std::to_string(prop.minor));
    DevPlan d;
    d.num_sms = prop.multiProcessorCount;
    CUDA_TRY(setup_schedule(hp->p.fwd, false, d.num_sms, d.fwd));
    CUDA_TRY(setup_schedule(hp->p.bwd, true, d.num_sms, d.bwd));
    CUDA_TRY(upload(&d.pack_src, hp->p.pack_src));
    CUDA_TRY(upload(&d.unpack_src, hp->p.unpack_src));
    if (hp->mma.ok) {
        CUDA_TRY(mma_setup(hp->mma, d.num_sms, d.mma));
    }
    if (hp->chain.ok) {
        CUDA_TRY(chain_setup(hp->p, hp->chain, d.num_sms, d.chain));
    }
    if (hp->tc3.ok) {
        CUDA_TRY(tc3_setup(hp->tc3, d.num_sms, d.tc3));
    }
    if (hp->tc3f.ok && hp->tc3i.ok) {
        CUDA_TRY(tc3_setup(hp->tc3f, d.num_sms, d.tc3f));
        CUDA_TRY(tc3_setup(hp->tc3i, d.num_sms, d.tc3i));
    }
    auto res = hp->dev.emplace(dev, d);
    *out = &res.first->second;
    return HINT_OK;
}

DevTables make_tables(const Plan& p, const Schedule& s, const DevSchedule& ds) {
    DevTables t;
    t.cgs = ds.cgs; t.eps = ds.eps; t.dwjobs = ds.dwjobs; t.stages = ds.stages;
    t.nstages = (int)s.stages.size();
    t.d = p.d; t.dc = p.dc;
    t.col_x = s.col_x; t.col_d = s.col_d; t.col_one = s.col_one; t.col_zero = s.col_zero;
    t.raw_off = s.raw_off;
    t.alpha = p.alpha;
    return t;
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int pack_weights(const hint_plan* hp, const DevPlan& d, const float* params, float* packed, cudaStream_t st) {
    const long long n = hp->p.n_packed;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
    hint_pack_kernel<<<blocks, threads, 0, st>>>(d.pack_src, params, packed, n); HINT_LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return HINT_OK;
}

long long bwd_ctas(const Plan& p, const DevPlan& d, long long B) {
    const long long ntiles = (B + p.bwd.TM - 1) / p.bwd.TM;
    return std::min<long long>(ntiles, d.bwd.max_ctas);
}

long long mma_bwd_ctas(const MmaPlan& m, const DevPlan& d, long long B) {
    const long long ntiles = (B + m.bwd.TM - 1) / m.bwd.TM;
    return std::min<long long>(ntiles, d.mma.bwd.max_ctas);
}

// packed operands of the warp-MMA path: [hi | lo], each n_packed floats (256-byte aligned)
size_t mma_half_bytes(const MmaPlan& m) { return align256((size_t)mma_copy_stride(m) * mma_weight_copies() * 4); }
size_t mma_packed_bytes(const MmaPlan& m) { return 2 * mma_half_bytes(m); }

}  // namespace

extern "C" {

const char* hint_last_error(void) { return g_err.c_str(); }

// Developer entry point: the regular tc3 backward launch with a timeline of CTA 0's second tile written to `prof`
// (long long[4096]: [0,1024) issue clock of every MMA record, [1024,2048) arrival clocks of the epilogue steps of warps 0 / 4,
// [2048, ...) clock at which warp 0 passed the step's wait).  Read by tests/cuda/prof_tc3.py.
int hint_dev_tc3_profile(hint_plan_t* hp, const float* z, const float* c, const float* params, const float* dz, const float* dlogdet,
                         int64_t B, float* dx, float* dc, float* dparams, long long* prof, int32_t* info, void* workspace, void* stream) {
    if (!hp || !hp->tc3.ok) return fail(HINT_ERR_UNSUPPORTED, "no tc3 plan");
    DevPlan* d = nullptr;
    int rc = get_dev(hp, &d);
    if (rc != HINT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* packed = reinterpret_cast<float*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((hp->tc3.n_packed * 4 + 255) & ~255ll));
    CUDA_TRY(tc3_pack(hp->tc3, d->tc3, params, packed, st));
    const int grid = tc3_bwd_ctas(d->tc3, (long long)B);
    CUDA_TRY(tc3_launch_bwd(hp->tc3, d->tc3, grid, z, c, packed, dz, dlogdet, nullptr, dx, dc, partials, dparams, (long long)B, st, prof));
    const T3Plan& t = hp->tc3;
    info[0] = (int32_t)t.mmas.size(); info[1] = (int32_t)t.epis.size(); info[2] = (int32_t)t.n_mma_instr; info[3] = (int32_t)t.tensor_cycles;
    for (size_t i = 0; i < t.epis.size() && i < 512; ++i) info[16 + i] = t.epis[i].type | (t.epis[i].wait_mma << 8);
    for (size_t i = 0; i < t.mmas.size() && i < 1024; ++i) info[1024 + i] = (int32_t)t.mmas[i].nk | ((int32_t)t.mmas[i].flags << 8) | ((int32_t)(((t.mmas[i].idesc >> 17) & 63) << 3) << 16);
    return HINT_OK;
}

// Developer entry point (not part of include/hint_b200.h): step-limited single-tile run of the tcgen05 training kernel with a
// dump of TMEM and shared memory, compared step by step against tests/emul/emul_tc3.cpp by tests/cuda/dbg_tc3.py.
int hint_dev_tc3_debug(hint_plan_t* hp, int32_t n_epi_limit, const float* z, const float* c, const float* params, const float* dz,
                       const float* dlogdet, int64_t B, float* x_rec, float* dx, float* dc, float* dump, int32_t* layout, void* workspace,
                       void* stream) {
    if (!hp || !hp->tc3.ok) return fail(HINT_ERR_UNSUPPORTED, "no tc3 plan");
    DevPlan* d = nullptr;
    int rc = get_dev(hp, &d);
    if (rc != HINT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* packed = reinterpret_cast<float*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((hp->tc3.n_packed * 4 + 255) & ~255ll));
    CUDA_TRY(tc3_pack(hp->tc3, d->tc3, params, packed, st));
    CUDA_TRY(tc3_debug_run(hp->tc3, d->tc3, n_epi_limit, z, c, packed, dz, dlogdet, x_rec, dx, dc, partials, (long long)B, dump, st));
    const T3Plan& t = hp->tc3;
    int32_t L[] = {t.sm_xs, t.sm_gs, t.sm_os, t.xp, t.op, t.sm_img[0], t.sm_img[1], t.sm_img[2], t.sm_img[3], t.sm_img[4],
                   t.img_rows[0], t.img_rows[1], t.img_rows[2], t.img_rows[3], t.img_rows[4], (int32_t)t.epis.size(), t.sm_red};
    for (size_t i = 0; i < sizeof(L) / 4; ++i) layout[i] = L[i];
    return HINT_OK;
}
const char* hint_version(void) { return "hint_b200 0.2 sm_100a"; }

uint64_t hint_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }

int hint_plan_create(int32_t d, int32_t dc, const int32_t* c_internal, int32_t n_internal, double clamp,
                     int32_t max_splits, int32_t min_split_size, int32_t reshuffle, hint_plan_t** out) {
    if (!out) return fail(HINT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    hint_plan* hp = new hint_plan();
    int code = HINT_OK;
    std::string err = build_plan(hp->p, d, dc, c_internal, n_internal, clamp, max_splits, min_split_size, reshuffle, &code);
    if (!err.empty()) {
        delete hp;
        return fail(code, err);
    }
    build_mma_plan(hp->p, hp->mma);
    build_chain_plan(hp->p, hp->chain);
    if (hp->chain.ok && !chain_fits(hp->p, hp->chain, &hp->chain.why)) hp->chain.ok = false;
    build_tc3_plan(hp->p, hp->tc3);
    build_tc3_plan(hp->p, hp->tc3f, T3K_FORWARD);
    build_tc3_plan(hp->p, hp->tc3i, T3K_INVERSE);
    *out = hp;
    return HINT_OK;
}

void hint_plan_destroy(hint_plan_t* hp) {
    if (!hp) return;
    for (auto& kv : hp->dev) {
        DevPlan& d = kv.second;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) break;
        if (cudaSetDevice(kv.first) != cudaSuccess) continue;
        for (DevSchedule* s : {&d.fwd, &d.bwd}) {
            cudaFree(s->cgs); cudaFree(s->eps); cudaFree(s->dwjobs); cudaFree(s->stages);
        }
        cudaFree(d.pack_src); cudaFree(d.unpack_src);
        mma_free(d.mma);
        chain_free(d.chain);
        tc3_free(d.tc3);
        tc3_free(d.tc3f);
        tc3_free(d.tc3i);
        cudaSetDevice(cur);
    }
    delete hp;
}

int32_t hint_plan_num_nodes(const hint_plan_t* hp) { return hp ? (int32_t)hp->p.nodes.size() : 0; }

int hint_plan_node(const hint_plan_t* hp, int32_t idx, hint_node_info_t* out) {
    if (!hp || !out || idx < 0 || idx >= (int32_t)hp->p.nodes.size()) return fail(HINT_ERR_INVALID, "bad node index");
    *out = hp->p.nodes[idx];
    return HINT_OK;
}

int64_t hint_plan_param_count(const hint_plan_t* hp) { return hp ? hp->p.n_params : 0; }

int hint_plan_param_layout(const hint_plan_t* hp, int64_t* offsets, int64_t n_offsets) {
    if (!hp || !offsets || n_offsets != (int64_t)hp->p.param_offsets.size())
        return fail(HINT_ERR_INVALID, "offsets must hold num_nodes*12 entries");
    std::memcpy(offsets, hp->p.param_offsets.data(), sizeof(int64_t) * (size_t)n_offsets);
    return HINT_OK;
}

int64_t hint_plan_flops_per_sample(const hint_plan_t* hp) { return hp ? hp->p.flops : 0; }

int32_t hint_plan_tile_rows(const hint_plan_t* hp, int32_t which) {
    if (!hp) return 0;
    return which == HINT_WS_BACKWARD ? hp->p.bwd.TM : hp->p.fwd.TM;
}

int32_t hint_plan_mode_supported(const hint_plan_t* hp, int32_t mode) {
    if (!hp) return 0;
    switch (mode) {
        case HINT_MODE_FP32: return 1;
        case HINT_MODE_TF32: case HINT_MODE_TF32X3: case HINT_MODE_TF32_MMA: return hp->mma.ok ? 1 : 0;
        case HINT_MODE_TF32_CHAIN: return (hp->chain.ok && hp->mma.ok) ? 1 : 0;
        case HINT_MODE_TF32_TC3: case HINT_MODE_TF32_TCGEN05: return (hp->tc3.ok && hp->mma.ok) ? 1 : 0;
    }
    return 0;
}

size_t hint_workspace_bytes(const hint_plan_t* hp_c, int64_t B, int32_t which) {
    hint_plan* hp = const_cast<hint_plan*>(hp_c);
    if (!hp || B < 0) { fail(HINT_ERR_INVALID, "bad plan or batch"); return 0; }
    size_t bytes = align256((size_t)hp->p.n_packed * 4);
    if (hp->mma.ok) bytes = std::max(bytes, mma_packed_bytes(hp->mma));
    if (hp->chain.ok) bytes = std::max(bytes, align256((size_t)hp->chain.n_packed * 4));
    if (hp->tc3.ok) bytes = std::max(bytes, align256((size_t)hp->tc3.n_packed * 4));
    if (hp->tc3f.ok) bytes = std::max(bytes, align256((size_t)std::max(hp->tc3f.n_packed, hp->tc3i.n_packed) * 4));
    if (which == HINT_WS_BACKWARD) {
        DevPlan* d = nullptr;
        if (get_dev(hp, &d) != HINT_OK) return 0;
        size_t part = align256((size_t)bwd_ctas(hp->p, *d, B) * (size_t)hp->p.n_partial * 4);
        if (hp->mma.ok) part = std::max(part, align256((size_t)mma_bwd_ctas(hp->mma, *d, B) * (size_t)hp->mma.n_partial * 4));
        if (hp->chain.ok) part = std::max(part, align256((size_t)chain_bwd_ctas(hp->chain, d->chain, B) * (size_t)hp->chain.n_partial * 4));
        if (hp->tc3.ok) part = std::max(part, align256((size_t)tc3_bwd_ctas(d->tc3, B) * (size_t)hp->tc3.n_partial * 4));
        bytes += part;
    }
    return bytes + 256;
}

static int check_common(const hint_plan* hp, const float* x, const float* c, const float* params, int64_t B, int32_t mode) {
    if (!hp) return fail(HINT_ERR_INVALID, "plan is NULL");
    if (B < 0) return fail(HINT_ERR_INVALID, "negative batch");
    if (mode != HINT_MODE_FP32 && mode != HINT_MODE_TF32 && mode != HINT_MODE_TF32X3 && mode != HINT_MODE_TF32_TCGEN05 &&
        mode != HINT_MODE_TF32_MMA && mode != HINT_MODE_TF32_CHAIN && mode != HINT_MODE_TF32_TC3)
        return fail(HINT_ERR_INVALID, "unknown mode");
    if (mode == HINT_MODE_TF32_TCGEN05) mode = HINT_MODE_TF32_TC3;   // one tcgen05 / TMEM machine: the older name is an alias
    if (mode == HINT_MODE_TF32_TC3 && !hp->tc3.ok)
        return fail(HINT_ERR_UNSUPPORTED, "this block is outside the tcgen05 kernels' envelope: " + hp->tc3.why);
    if (mode == HINT_MODE_TF32_TC3) mode = HINT_MODE_TF32;   // forward / inverse: the TF32 default
    if (mode == HINT_MODE_TF32_CHAIN && !hp->chain.ok)
        return fail(HINT_ERR_UNSUPPORTED, "this block is outside the register-chained kernels' envelope: " + hp->chain.why);
    if (mode == HINT_MODE_TF32_CHAIN) mode = HINT_MODE_TF32_MMA;   // same requirements otherwise
    if ((mode == HINT_MODE_TF32 || mode == HINT_MODE_TF32X3 || mode == HINT_MODE_TF32_MMA) && !hp->mma.ok)
        return fail(HINT_ERR_UNSUPPORTED, "this block is outside the warp-MMA kernels' envelope: " + hp->mma.why);
    if (B > 0 && (!x || !params)) return fail(HINT_ERR_INVALID, "NULL input pointer");
    if (B > 0 && hp->p.dc > 0 && !c) return fail(HINT_ERR_INVALID, "plan has a condition input but c is NULL");
    if (!aligned16(x) || !aligned16(c) || !aligned16(params)) return fail(HINT_ERR_INVALID, "pointers must be 16-byte aligned");
    return HINT_OK;
}

int hint_forward(const hint_plan_t* hp_c, const float* x, const float* c, const float* params, int64_t B, int32_t rev,
                 int32_t mode, float* z, float* logdet, void* workspace, size_t workspace_bytes, void* stream) {
    hint_plan* hp = const_cast<hint_plan*>(hp_c);
    int rc = check_common(hp, x, c, params, B, mode);
    if (rc != HINT_OK) return rc;
    if (B == 0) return HINT_OK;
    if (!z || !logdet) return fail(HINT_ERR_INVALID, "NULL output pointer");
    if (!aligned16(z) || !aligned16(workspace)) return fail(HINT_ERR_INVALID, "pointers must be 16-byte aligned");
    if (z == x) return fail(HINT_ERR_INVALID, "z must not alias x");
    DevPlan* d = nullptr;
    if ((rc = get_dev(hp, &d)) != HINT_OK) return rc;
    if (!workspace || workspace_bytes < hint_workspace_bytes(hp, B, HINT_WS_FORWARD))
        return fail(HINT_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* packed = reinterpret_cast<float*>(workspace);
    // HINT_MODE_TF32 forward / inverse: two tensor-core kernels exist.  The tcgen05/TMEM kernel is the faster one where the
    // block fits its TMEM envelope (measured 1.9 ms vs 4.0 ms per 2^20 samples on the d=43 hint_8 block), the warp-MMA kernel
    // covers the rest.  HINT_B200_TF32_FWD=mma|tcgen05 forces one; HINT_MODE_TF32_MMA / HINT_MODE_TF32_TCGEN05 name them
    // explicitly (tests run both).
    // The tcgen05 / TMEM machine of the training kernel also runs the transport alone (T3K_FORWARD / T3K_INVERSE programs):
    // explicit with HINT_MODE_TF32_TC3, and the HINT_MODE_TF32 default for blocks the register-chained kernels do not cover
    // (measured on the gas block: 2.5x the older tcgen05 forward kernel, which keeps the blocks outside this envelope).
    if (mode == HINT_MODE_TF32_TCGEN05) mode = HINT_MODE_TF32_TC3;
    const bool tc3_transport = hp->tc3f.ok && hp->tc3i.ok &&
                               (mode == HINT_MODE_TF32_TC3 || (mode == HINT_MODE_TF32 && !hp->chain.ok && !dev_getenv("HINT_B200_TF32_FWD")));
    if (tc3_transport) {
        const T3Plan& t = rev ? hp->tc3i : hp->tc3f;
        const DevTc3& dv = rev ? d->tc3i : d->tc3f;
        CUDA_TRY(tc3_pack(t, dv, params, packed, st));
        CUDA_TRY(tc3_launch_transport(t, dv, x, c, packed, z, logdet, (long long)B, st));
        return HINT_OK;
    }
    if (mode == HINT_MODE_TF32_TC3) mode = HINT_MODE_TF32;   // outside the transport envelope: forward / inverse as in HINT_MODE_TF32
    if (mode == HINT_MODE_TF32) {
        static const char* pref = dev_getenv("HINT_B200_TF32_FWD");
        const bool want_chain = pref ? std::strcmp(pref, "chain") == 0 : true;
        mode = (want_chain && hp->chain.ok) ? HINT_MODE_TF32_CHAIN : HINT_MODE_TF32_MMA;
    }
    if (mode == HINT_MODE_TF32_CHAIN) {
        CUDA_TRY(chain_pack(hp->chain, d->chain, params, packed, st));
        CUDA_TRY(chain_launch_fwd(hp->p, hp->chain, d->chain, x, c, packed, z, logdet, (long long)B, rev ? 1 : 0, st));
        return HINT_OK;
    }
    if (mode == HINT_MODE_TF32_MMA || mode == HINT_MODE_TF32X3) {
        const bool x3 = mode == HINT_MODE_TF32X3;
        float* hi = packed;
        float* lo = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mma_half_bytes(hp->mma));
        CUDA_TRY(mma_pack(hp->mma, d->mma, params, hi, x3 ? lo : nullptr, st));
        CUDA_TRY(mma_launch_fwd(hp->p, hp->mma, d->mma, x3, x, c, hi, lo, z, logdet, (long long)B, rev ? 1 : 0, st));
        return HINT_OK;
    }
    if ((rc = pack_weights(hp, *d, params, packed, st)) != HINT_OK) return rc;
    const Schedule& s = hp->p.fwd;
    const DevTables T = make_tables(hp->p, s, d->fwd);
    const long long ntiles = (B + s.TM - 1) / s.TM;
    const int grid = (int)std::min<long long>(ntiles, d->fwd.max_ctas);
#define LAUNCH_FWD(TMV)                                                                                   \
    case TMV:                                                                                             \
        hint_fwd_fp32_kernel<TMV><<<grid, kThreads, s.smem_bytes, st>>>(T, x, c, packed, z, logdet, (long long)B, rev ? 1 : 0); HINT_LAUNCHED(); \
        break;
    switch (s.TM) {
        LAUNCH_FWD(128) LAUNCH_FWD(64) LAUNCH_FWD(32) LAUNCH_FWD(16) LAUNCH_FWD(8)
        default: return fail(HINT_ERR_INVALID, "bad tile size");
    }
#undef LAUNCH_FWD
    CUDA_TRY(cudaGetLastError());
    return HINT_OK;
}

// nll_scale != 0 (with dz == dlogdet == NULL): the upstream gradient is generated in the tile load, dz = nll_scale * z and
// dlogdet = -nll_scale (hint_backward_nll); only the kernel families that implement it accept that form.
static int backward_impl(const hint_plan_t* hp_c, const float* z, const float* c, const float* params, const float* dz,
                         const float* dlogdet, float nll_scale, int64_t B, int32_t mode, float* x_rec, float* dx, float* dc, float* dparams,
                         void* workspace, size_t workspace_bytes, void* stream) {
    hint_plan* hp = const_cast<hint_plan*>(hp_c);
    if (mode == HINT_MODE_TF32_TCGEN05) mode = HINT_MODE_TF32_TC3;   // alias
    int rc = check_common(hp, z, c, params, B, mode);
    if (rc != HINT_OK) return rc;
    if (!dparams) return fail(HINT_ERR_INVALID, "dparams is NULL");
    // HINT_MODE_TF32 backward: the register-chained kernel where the block fits its shape table, else the interpreter
    // warp-MMA kernel.  HINT_B200_TF32_BWD=mma forces the latter (developer aid; HINT_MODE_TF32_MMA does the same).
    bool use_chain = mode == HINT_MODE_TF32_CHAIN;
    if (mode == HINT_MODE_TF32 && hp->chain.ok) {
        static const char* pref = dev_getenv("HINT_B200_TF32_BWD");
        use_chain = !(pref && std::strcmp(pref, "mma") == 0);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        CUDA_TRY(cudaMemsetAsync(dparams, 0, (size_t)hp->p.n_params * 4, st));
        return HINT_OK;
    }
    const bool nll = nll_scale != 0.f;
    if ((!nll && (!dz || !dlogdet)) || !dx) return fail(HINT_ERR_INVALID, "NULL gradient pointer");
    if (nll && !(mode == HINT_MODE_TF32_TC3 || mode == HINT_MODE_TF32_CHAIN || (mode == HINT_MODE_TF32 && (use_chain || hp->tc3.ok))))
        return fail(HINT_ERR_UNSUPPORTED, "the fused NLL gradient exists in the register-chained and tcgen05 training kernels only");
    if (!aligned16(dz) || !aligned16(dx) || !aligned16(dc) || !aligned16(x_rec) || !aligned16(workspace))
        return fail(HINT_ERR_INVALID, "pointers must be 16-byte aligned");
    DevPlan* d = nullptr;
    if ((rc = get_dev(hp, &d)) != HINT_OK) return rc;
    if (!workspace || workspace_bytes < hint_workspace_bytes(hp, B, HINT_WS_BACKWARD))
        return fail(HINT_ERR_WORKSPACE, "workspace too small");
    float* packed = reinterpret_cast<float*>(workspace);
    // tcgen05 training kernel: explicit mode, and the HINT_MODE_TF32 default for blocks the register-chained kernels do not cover
    if (mode == HINT_MODE_TF32_TC3 || (mode == HINT_MODE_TF32 && !use_chain && hp->tc3.ok)) {
        float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256((size_t)hp->tc3.n_packed * 4));
        CUDA_TRY(tc3_pack(hp->tc3, d->tc3, params, packed, st));
        const int grid = tc3_bwd_ctas(d->tc3, (long long)B);
        CUDA_TRY(tc3_launch_bwd(hp->tc3, d->tc3, grid, z, c, packed, dz, dlogdet, x_rec, dx, dc, partials, dparams, (long long)B, st, nullptr, nll_scale));
        return HINT_OK;
    }
    if (use_chain) {
        float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256((size_t)hp->chain.n_packed * 4));
        CUDA_TRY(chain_pack(hp->chain, d->chain, params, packed, st));
        const int grid = chain_bwd_ctas(hp->chain, d->chain, (long long)B);
        CUDA_TRY(chain_launch_bwd(hp->p, hp->chain, d->chain, grid, z, c, packed, dz, dlogdet, x_rec, dx, dc, partials, (long long)B, st, nll_scale));
        const long long n = hp->p.n_params;
        const int threads = 256;
        const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
        hint_reduce_unpack_kernel<<<blocks, threads, 0, st>>>(d->chain.unpack_src, partials, grid, (long long)hp->chain.n_partial, dparams, n); HINT_LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        return HINT_OK;
    }
    if (mode == HINT_MODE_TF32 || mode == HINT_MODE_TF32X3 || mode == HINT_MODE_TF32_MMA || mode == HINT_MODE_TF32_CHAIN) {
        const bool x3 = mode == HINT_MODE_TF32X3;
        float* hi = packed;
        float* lo = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mma_half_bytes(hp->mma));
        float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mma_packed_bytes(hp->mma));
        CUDA_TRY(mma_pack(hp->mma, d->mma, params, hi, x3 ? lo : nullptr, st));
        const int grid = (int)mma_bwd_ctas(hp->mma, *d, B);
        const long long np = hp->mma.n_partial;
        CUDA_TRY(mma_launch_bwd(hp->p, hp->mma, d->mma, x3, grid, z, c, hi, lo, dz, dlogdet, x_rec, dx, dc, partials, (long long)B, st));
        const long long n = hp->p.n_params;
        const int threads = 256;
        const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
        hint_reduce_unpack_kernel<<<blocks, threads, 0, st>>>(d->mma.unpack_src, partials, grid, np, dparams, n); HINT_LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        return HINT_OK;
    }
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256((size_t)hp->p.n_packed * 4));
    if ((rc = pack_weights(hp, *d, params, packed, st)) != HINT_OK) return rc;
    const Schedule& s = hp->p.bwd;
    const DevTables T = make_tables(hp->p, s, d->bwd);
    const int grid = (int)bwd_ctas(hp->p, *d, B);
#define LAUNCH_BWD(TMV)                                                                                        \
    case TMV:                                                                                                  \
        hint_bwd_fp32_kernel<TMV><<<grid, kThreads, s.smem_bytes, st>>>(T, z, c, packed, dz, dlogdet, x_rec, dx, dc, partials, \
                                                                         (long long)hp->p.n_partial, (long long)B); HINT_LAUNCHED(); \
        break;
    switch (s.TM) {
        LAUNCH_BWD(128) LAUNCH_BWD(64) LAUNCH_BWD(32) LAUNCH_BWD(16) LAUNCH_BWD(8)
        default: return fail(HINT_ERR_INVALID, "bad tile size");
    }
#undef LAUNCH_BWD
    CUDA_TRY(cudaGetLastError());
    {
        const long long n = hp->p.n_params;
        const int threads = 256;
        const int blocks = (int)std::min<long long>((n + threads - 1) / threads, 148 * 8);
        hint_reduce_unpack_kernel<<<blocks, threads, 0, st>>>(d->unpack_src, partials, grid, (long long)hp->p.n_partial, dparams, n); HINT_LAUNCHED();
        CUDA_TRY(cudaGetLastError());
    }
    return HINT_OK;
}

int hint_backward(const hint_plan_t* hp, const float* z, const float* c, const float* params, const float* dz,
                  const float* dlogdet, int64_t B, int32_t mode, float* x_rec, float* dx, float* dc, float* dparams,
                  void* workspace, size_t workspace_bytes, void* stream) {
    if (B > 0 && (!dz || !dlogdet)) return fail(HINT_ERR_INVALID, "NULL gradient pointer");
    return backward_impl(hp, z, c, params, dz, dlogdet, 0.f, B, mode, x_rec, dx, dc, dparams, workspace, workspace_bytes, stream);
}

int hint_backward_nll(const hint_plan_t* hp, const float* z, const float* c, const float* params, const float* dz, float grad_scale,
                      int64_t B, int32_t mode, float* x_rec, float* dx, float* dc, float* dparams, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!(grad_scale != 0.f)) return fail(HINT_ERR_INVALID, "grad_scale must be non-zero (1/B for the mean NLL)");
    return backward_impl(hp, z, c, params, dz, nullptr, grad_scale, B, mode, x_rec, dx, dc, dparams, workspace, workspace_bytes, stream);
}

int hint_add_noise(const float* x, float* out, int64_t n, float sigma, uint64_t seed, uint64_t offset, void* stream) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(HINT_ERR_INVALID, "bad noise arguments");
    CUDA_TRY(train_add_noise(x, out, (long long)n, sigma, seed, offset, (cudaStream_t)stream));
    return HINT_OK;
}

size_t hint_nll_workspace_bytes(void) { return train_nll_workspace_bytes(); }

int hint_nll_loss(const float* z, const float* const* logdets, int32_t n_logdets, int64_t B, int32_t d, float* loss3, void* workspace,
                  size_t workspace_bytes, void* stream) {
    if (B <= 0 || d <= 0 || !z || !logdets || n_logdets < 1 || n_logdets > 64 || !loss3) return fail(HINT_ERR_INVALID, "bad NLL arguments");
    if (!workspace || workspace_bytes < train_nll_workspace_bytes()) return fail(HINT_ERR_WORKSPACE, "workspace too small");
    CUDA_TRY(train_nll_loss(z, logdets, n_logdets, (long long)B, d, loss3, workspace, (cudaStream_t)stream));
    return HINT_OK;
}

int hint_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float grad_clamp, int64_t step, void* stream) {
    if (n_tensors < 0 || (n_tensors > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes)))
        return fail(HINT_ERR_INVALID, "bad optimizer arguments");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t must be long long");
    CUDA_TRY(train_adam_step(n_tensors, params, grads, exp_avg, exp_avg_sq, reinterpret_cast<const long long*>(sizes), lr, beta1,
                             beta2, eps, weight_decay, grad_clamp, (long long)step, (cudaStream_t)stream));
    return HINT_OK;
}

int hint_householder_matrix(const float* Vs, int32_t n_reflections, int32_t d, float* W, void* stream) {
    if (!Vs || !W || d < 1 || d > hh_max_d() || n_reflections < 0) return fail(HINT_ERR_INVALID, "bad Householder arguments (1 <= d <= 128)");
    CUDA_TRY(hh_matrix(Vs, n_reflections, d, W, (cudaStream_t)stream));
    return HINT_OK;
}

int hint_householder_matrix_backward(const float* Vs, const float* W, const float* dW, int32_t n_reflections, int32_t d, float* dVs,
                                     void* stream) {
    if (!Vs || !W || !dW || !dVs || d < 1 || d > hh_max_d() || n_reflections < 0)
        return fail(HINT_ERR_INVALID, "bad Householder arguments (1 <= d <= 128)");
    CUDA_TRY(hh_matrix_backward(Vs, W, dW, n_reflections, d, dVs, (cudaStream_t)stream));
    return HINT_OK;
}

int hint_householder_apply(const float* x, const float* W, int64_t B, int32_t d, int32_t transpose, float* y, void* stream) {
    if (B < 0 || d < 1 || d > hh_max_d() || (B > 0 && (!x || !W || !y))) return fail(HINT_ERR_INVALID, "bad Householder arguments (1 <= d <= 128)");
    if (B > 0 && y == x) return fail(HINT_ERR_INVALID, "y must not alias x");
    CUDA_TRY(hh_apply(x, W, (long long)B, d, transpose, y, (cudaStream_t)stream));
    return HINT_OK;
}

size_t hint_householder_wgrad_workspace_bytes(int32_t d) { return d >= 1 && d <= hh_max_d() ? hh_wgrad_workspace_bytes(d) : 0; }

int hint_householder_wgrad(const float* x, const float* dz, int64_t B, int32_t d, float* dW, void* workspace, size_t workspace_bytes,
                           void* stream) {
    if (B < 0 || d < 1 || d > hh_max_d() || !dW || (B > 0 && (!x || !dz))) return fail(HINT_ERR_INVALID, "bad Householder arguments (1 <= d <= 128)");
    if (!workspace || workspace_bytes < hh_wgrad_workspace_bytes(d)) return fail(HINT_ERR_WORKSPACE, "workspace too small");
    CUDA_TRY(hh_wgrad(x, dz, (long long)B, d, dW, workspace, (cudaStream_t)stream));
    return HINT_OK;
}

int hint_mlp_coupling_supported(int32_t du, int32_t dv, int32_t hidden) { return mc_supported(du, dv, hidden) ? 1 : 0; }

int hint_mlp_coupling_forward(const float* u, int32_t du, const float* v, int32_t dv, int32_t hidden, const float* const* params, float clamp,
                              int32_t rev, int64_t B, float* y, float* logdet, void* stream) {
    if (!mc_supported(du, dv, hidden)) return fail(HINT_ERR_UNSUPPORTED, "coupling outside the fused kernel's envelope (du, dv <= 128, hidden <= 256)");
    if (B < 0 || !params || (B > 0 && (!u || !v || !y || !logdet))) return fail(HINT_ERR_INVALID, "bad coupling arguments");
    for (int i = 0; i < 16; ++i) if (!params[i]) return fail(HINT_ERR_INVALID, "null parameter pointer");
    if (B > 0 && y == v) return fail(HINT_ERR_INVALID, "y must not alias v");
    CUDA_TRY(mc_forward(u, du, v, dv, hidden, params, clamp, rev ? 1 : 0, (long long)B, y, logdet, (cudaStream_t)stream));
    return HINT_OK;
}

size_t hint_mlp_coupling_workspace_bytes(int32_t du, int32_t dv, int32_t hidden, int64_t B) {
    return mc_supported(du, dv, hidden) && B >= 0 ? mc_workspace_bytes(du, dv, hidden, (long long)B) : 0;
}

int hint_mlp_coupling_backward(const float* u, int32_t du, const float* v, int32_t dv, int32_t hidden, const float* const* params, float clamp,
                               int64_t B, const float* dy, const float* dlogdet, float* du_grad, float* dv_grad, float* const* dparams,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (!mc_supported(du, dv, hidden)) return fail(HINT_ERR_UNSUPPORTED, "coupling outside the fused kernel's envelope (du, dv <= 128, hidden <= 256)");
    if (B < 0 || !params || !dparams || (B > 0 && (!u || !v || !dy || !du_grad || !dv_grad))) return fail(HINT_ERR_INVALID, "bad coupling arguments");
    for (int i = 0; i < 16; ++i) if (!params[i] || !dparams[i]) return fail(HINT_ERR_INVALID, "null parameter pointer");
    if (!workspace || workspace_bytes < mc_workspace_bytes(du, dv, hidden, (long long)B)) return fail(HINT_ERR_WORKSPACE, "workspace too small");
    CUDA_TRY(mc_backward(u, du, v, dv, hidden, params, clamp, (long long)B, dy, dlogdet, du_grad, dv_grad, dparams, workspace, workspace_bytes,
                         (cudaStream_t)stream));
    return HINT_OK;
}

size_t hint_mmd_workspace_bytes(int64_t n) { return n >= 1 ? mmd_workspace_bytes((long long)n) : 0; }

int hint_multi_mmd(const float* x, const float* y, int64_t n, int32_t d, const float* widths, const float* exponents, int32_t n_kernels,
                   float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 1 || d < 1 || !x || !y || !out || !widths || !exponents || n_kernels < 1 || n_kernels > 8)
        return fail(HINT_ERR_INVALID, "bad MMD arguments (n >= 1, d >= 1, 1 <= n_kernels <= 8)");
    for (int i = 0; i < n_kernels; ++i)
        if (!(widths[i] > 0.f) || !(exponents[i] > 0.f)) return fail(HINT_ERR_INVALID, "MMD kernel widths and exponents must be positive");
    if (!workspace || workspace_bytes < mmd_workspace_bytes((long long)n)) return fail(HINT_ERR_WORKSPACE, "workspace too small");
    CUDA_TRY(mmd_multi(x, y, (long long)n, d, widths, exponents, n_kernels, out, workspace, workspace_bytes, (cudaStream_t)stream));
    return HINT_OK;
}

}  // extern "C"
