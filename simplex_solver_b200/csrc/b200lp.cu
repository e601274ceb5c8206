// b200lp.cu -- host side of libb200lp.so: the C ABI of include/b200lp.h over the sm_100a kernels.
//
// The pivot loop is device resident; the kernels communicate through a DevState in device memory and the host only
// reads the status word back between CUDA-graph replays (double buffered, so the GPU never waits for the host).
// Three drivers (b200lp_opts.loop_mode):
//   rank-1 graph loop : per pivot [k_pick_cluster (or k_price + k_ratio)] -> k_update_ldg / k_update_tma
//   on-chip loop      : one persistent cooperative kernel, tableau resident in shared memory (k_solve_onchip)
//   look-ahead loop   : K x k_pick_cluster<.., BLOCKED> -> k_blk_row -> k_blk_flush_special + k_blk_flush_db
//                       (one tableau pass per K pivots)
// No cuBLAS, no Triton, no CPU fallback: without a CUDA device every compute entry point fails with
// B200LP_E_CUDA.
#include "../../include/b200lp.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_batched.cuh"
#include "kernels_blocked.cuh"
#include "kernels_build.cuh"
#include "kernels_cluster.cuh"
#include "kernels_flush_mma.cuh"
#include "kernels_onchip.cuh"
#include "kernels_pick.cuh"
#include "kernels_picks.cuh"
#include "kernels_shard.cuh"
#include "kernels_update.cuh"

using namespace b200lp;

#define B200LP_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(B200LP_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define CKR(call)               \
    do {                        \
        int rc__ = (call);      \
        if (rc__) return rc__;  \
    } while (0)

static const size_t BATCHED_SMEM_MAX = 200 * 1024;  // dynamic shared memory the batched kernel may opt in to
static const int BATCHED_MAX_CHUNKS = 16;           // chunks of a host-side batch (and work counters of their launches)

// ------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------
// Guard mode (environment B200LP_GUARD=1, read once): every workspace buffer is allocated between two 4 KB bands
// filled with a byte pattern, and b200lp_check_guards() counts the band bytes that no longer hold it.  It stands in for
// compute-sanitizer's memcheck on boxes that refuse the tool: an out-of-bounds store next to any library-owned buffer
// (pivot column copy, look-ahead history, partials, staging) shows up in the tests instead of corrupting a neighbour.
static const size_t GUARD_BYTES = 4096;
static const int GUARD_PATTERN = 0xA5;
static bool guard_mode() {
    static const bool on = [] {
        const char* e = getenv("B200LP_GUARD");
        return e && *e && *e != '0';
    }();
    return on;
}

// Bumped whenever a workspace buffer is (re)allocated or released, by any solver of the process: launches that a CALLER
// captured into a CUDA graph (simplex_solver_b200/sharded.py) bake device pointers and capacities, so a cached graph is
// only valid for the epoch it was captured in (b200lp_binding_epoch).
static std::atomic<long long> g_alloc_epoch{0};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    bool guarded = false;
    void* base() const { return guarded ? (void*)((char*)p - GUARD_BYTES) : (void*)p; }
    int ensure(size_t n) {
        if (n <= cap) return 0;
        release();
        g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
        const bool g = guard_mode();
        const size_t payload = (n * sizeof(T) + 255) / 256 * 256;
        void* raw = nullptr;
        cudaError_t e = cudaMalloc(&raw, payload + (g ? 2 * GUARD_BYTES : 0));
        if (e != cudaSuccess) return fail(B200LP_E_NOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
        if (g) {
            // (payload included: uninitialised reads become visible as pattern-valued doubles)
            // cudaMemset runs on the legacy default stream, which does not order with the solvers' non-blocking
            // streams: it must have finished before anybody enqueues work on the new buffer
            cudaMemset(raw, GUARD_PATTERN, payload + 2 * GUARD_BYTES);
            cudaStreamSynchronize(cudaStreamLegacy);
            raw = (char*)raw + GUARD_BYTES;
        }
        p = (T*)raw;
        guarded = g;
        cap = n;
        return 0;
    }
    // band bytes that were overwritten (0 when guard mode is off)
    long long check() const {
        if (!p || !guarded) return 0;
        const size_t payload = (cap * sizeof(T) + 255) / 256 * 256;
        std::vector<unsigned char> h(2 * GUARD_BYTES);
        if (cudaMemcpy(h.data(), (char*)p - GUARD_BYTES, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (cudaMemcpy(h.data() + GUARD_BYTES, (char*)p + payload, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        long long bad = 0;
        for (unsigned char c : h) bad += c != GUARD_PATTERN;
        return bad;
    }
    void release() {
        if (p) {
            cudaFree(base());
            g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
        }
        p = nullptr;
        cap = 0;
        guarded = false;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }  // b200lp_destroy deletes the solver: whatever it did not release by name goes here
};

// Everything a captured kernel launch bakes into the graph: the tableau binding, the options, and EVERY device pointer
// handed to a kernel of the loop.  A workspace is reused across unrelated problems (thread_solver), so two problems
// with the same R, C and tableau address but different m / art_base / label buffers must not share a graph.
struct GraphKey {
    const double* T = nullptr;
    int64_t R = 0, C = 0, ld = 0, obj_row = -1, m = -1;
    int32_t rule = -1, variant = -1, iters = 0, art_base = -1;
    double eps_cost = 0, eps_pivot = 0;
    int64_t hist_cap = 0;
    const double* snaps = nullptr;
    int64_t snap_cap = 0;
    int32_t blocked_k = 0;  // 0: rank-1 iterations; K > 0: look-ahead blocks of K pivots
    const void* ptrs[12] = {};  // rowlab, collab, col, history x 4, look-ahead buffers x 5
    bool operator==(const GraphKey& o) const {
        return blocked_k == o.blocked_k && snaps == o.snaps && snap_cap == o.snap_cap && T == o.T && R == o.R && C == o.C && ld == o.ld && obj_row == o.obj_row && rule == o.rule &&
               variant == o.variant && iters == o.iters && eps_cost == o.eps_cost && eps_pivot == o.eps_pivot &&
               hist_cap == o.hist_cap && m == o.m && art_base == o.art_base && memcmp(ptrs, o.ptrs, sizeof(ptrs)) == 0;
    }
};

struct b200lp_solver {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // second lane of the chunked host path of solve_batched
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_state[2] = {nullptr, nullptr};

    // tableau (attached: caller owned; otherwise `own_T`)
    double* T = nullptr;
    DevBuf<double> own_T;
    int64_t m = 0, n_obj = 0, R = 0, C = 0, ld = 0, n_struct = 0;
    int32_t art_base = 0;

    DevBuf<int32_t> rowlab, collab;
    DevBuf<double> col;
    DevBuf<Key> part_price, part_ratio;
    DevBuf<DevState> st;
    DevState* st_host = nullptr;  // pinned, 2 entries
    DevBuf<int32_t> h_row, h_col, h_enter, h_leave;
    int64_t hist_cap = 0;

    // staging for solve_dense / solve_batched
    DevBuf<double> sA, sb, sc, sx, sfun;
    DevBuf<RowInfo> sinfo;
    DevBuf<int8_t> sops;
    DevBuf<int32_t> sstatus, snpiv, slog;
    DevBuf<unsigned int> snext;  // work counters of the batched launches (one per chunk)

    // look-ahead (blocked) loop: pending pivots' columns / rows and the current objective row / right-hand side
    DevBuf<double> blk_colP, blk_qP, blk_obj, blk_rhs;
    DevBuf<BlkPending> blk_pend;
    BlkBuffers blk;

    // on-chip persistent loop: exchange buffer and grid-barrier counter
    DevBuf<double> xbuf;
    DevBuf<unsigned long long> gbar;

    // TMA descriptor of the current tableau
    CUtensorMap tmap;
    const double* tmap_T = nullptr;
    int64_t tmap_R = 0, tmap_C = 0, tmap_ld = 0;
    int tmap_box_r = 0;

    // optional dense copy of the tableau after every pivot (caller-owned device buffer)
    double* snaps = nullptr;
    int64_t snap_cap = 0;

    // peer-memory exchange of the sharded loops: the fused pick kernel reads its shard context from device memory
    // (one entry in production; the single-GPU emulation passes an array of them to ONE launch)
    P2PPeers p2p;
    bool p2p_on = false;
    DevBuf<long long> xgen;       // generation of the last completed exchange; zeroed by b200lp_p2p_connect only
    DevBuf<ShardCtx> shard_ctx;   // what k_shard_pick reads
    std::vector<ShardCtx> shard_ctx_host;  // last uploaded content (uploads happen only when something changed)

    // CTAs per cluster of the single-launch pick kernel (kernels_cluster.cuh); 0 = not available / switched off
    bool shard_la_fused = false;    // sharded look-ahead decision inside the one-cluster kernel (diagnostic)
    bool flush_mma = false;         // look-ahead flush through DMMA (k_blk_flush_mma) instead of the DFMA kernel
    int cluster_ctas = 0;
    bool coop_picks = false;        // look-ahead picks: one persistent cooperative kernel per block (kernels_picks.cuh)
    DevBuf<PickPartB> part_b;
    DevBuf<int32_t> pk_err;

    cudaGraphExec_t graph = nullptr;
    GraphKey graph_key;
    int64_t launches = 0;
    long long bind_epoch = 0;     // bumped by every (re)binding of the tableau, the snapshots and the peer regions

    // launch plan of the batched kernel, cached per LP shape (occupancy queries cost more than a small batch)
    struct BatchedPlan {
        int64_t m = -1, n = -1;
        int wpc = 1, per_sm = 0;
    } bplan;

    // wall-clock bound of the running call (b200lp_opts.time_limit_s): seconds on the steady clock, 0 = none
    double deadline = 0.0;
    double deadline_outer = 0.0;  // armed by b200lp_solve_dense so that the build counts, inherited by b200lp_solve
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void arm_deadline(b200lp_solver* s, const b200lp_opts* o) {
    if (s->deadline_outer > 0.0) s->deadline = s->deadline_outer;
    else s->deadline = o->time_limit_s > 0.0 ? now_s() + o->time_limit_s : 0.0;
}
static bool deadline_passed(const b200lp_solver* s) { return s->deadline > 0.0 && now_s() >= s->deadline; }

static int set_device(b200lp_solver* s) {
    CK(cudaSetDevice(s->device));
    return 0;
}

B200LP_API int b200lp_version(void) { return B200LP_VERSION; }
B200LP_API const char* b200lp_last_error(void) { return g_err.c_str(); }

B200LP_API void b200lp_default_opts(b200lp_opts* o) {
    if (!o) return;
    o->rule = B200LP_RULE_DANTZIG;
    o->update_variant = B200LP_UPDATE_AUTO;
    o->max_pivots = (int64_t)1 << 40;
    o->eps_cost = 1e-9;
    o->eps_pivot = 1e-9;
    o->eps_feas = 1e-7;
    o->check_every = 0;
    o->loop_mode = B200LP_LOOP_AUTO;
    o->time_limit_s = 0.0;
}

B200LP_API int b200lp_create(b200lp_solver** out, int device) {
    if (!out) return fail(B200LP_E_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200LP_E_CUDA, "no CUDA device (%s); libb200lp has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(B200LP_E_INVALID, "device %d out of range [0,%d)", device, count);
    b200lp_solver* s = new b200lp_solver();
    s->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    CK(cudaEventCreate(&s->ev0));
    CK(cudaEventCreate(&s->ev1));
    CK(cudaEventCreateWithFlags(&s->ev_state[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->ev_state[1], cudaEventDisableTiming));
    CK(cudaMallocHost(&s->st_host, 2 * sizeof(DevState)));
    CKR(s->st.ensure(1));
    CK(cudaMemsetAsync(s->st.p, 0, sizeof(DevState), s->stream));
    CKR(s->part_price.ensure(1024));
    CKR(s->part_ratio.ensure(1024));
    CKR(s->sfun.ensure(1));
    CK(cudaFuncSetAttribute(k_update_tma<TMA_BOX_R, TMA_STAGES, TMA_STORE_LAG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)TmaCfg<TMA_BOX_R, TMA_STAGES>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_solve_onchip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ONCHIP_SMEM_MAX));
    CK(cudaFuncSetAttribute(k_blk_flush_db<FL_WC, FL_TR, FL_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)FL_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_blk_flush_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_SMEM_BYTES));
    // diagnostic switch: the DMMA flush kernel.  Measured on B200 (scripts/probe_dmma.cu, scripts/probe_lookahead.py): every
    // f64 mma shape compiles to DMMA.8x8x4, which peaks at 12.3 T FMA/s against 16.0 T FMA/s for DFMA and shares its units
    // (mixed warps: 15.3 T in total), so the DMMA flush is bit-identical but slower: 23.6k vs 25.3k pivots/s at K = 32
    s->flush_mma = getenv("B200LP_FLUSH_MMA") != nullptr;
    s->shard_la_fused = getenv("B200LP_SHARD_LA_FUSED") != nullptr;
    if (!getenv("B200LP_NO_CLUSTER")) {  // diagnostic switch: fall back to the two-launch pick (k_price + k_ratio)
        const void* kernels[10] = {(const void*)k_pick_cluster<false, false>, (const void*)k_pick_cluster<true, false>,
                                   (const void*)k_pick_cluster<false, true>, (const void*)k_pick_cluster<true, true>,
                                   (const void*)k_shard_pick<false, false>, (const void*)k_shard_pick<true, false>,
                                   (const void*)k_shard_pick<false, true>, (const void*)k_shard_pick<true, true>,
                                   (const void*)k_shard_pick<false, true, true>, (const void*)k_shard_pick<true, true, true>};
        for (int nc : {16, 8}) {
            bool ok = true;
            for (const void* kf : kernels) {
                if (nc > 8 && cudaFuncSetAttribute(kf, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = false;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(nc);
                cfg.blockDim = dim3(CL_THREADS);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = nc;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                int n = 0;
                if (!ok || cudaOccupancyMaxActiveClusters(&n, kf, &cfg) != cudaSuccess || n < 1) ok = false;
            }
            cudaGetLastError();  // a failed probe must not poison later calls
            if (ok) {
                s->cluster_ctas = nc;
                break;
            }
        }
    }
    CK(cudaFuncSetAttribute(k_solve_batched<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCHED_SMEM_MAX));
    CK(cudaFuncSetAttribute(k_solve_batched<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCHED_SMEM_MAX));
    CK(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    CKR(s->snext.ensure(BATCHED_MAX_CHUNKS));
    CKR(s->gbar.ensure(1));
    {   // look-ahead picks as one persistent cooperative kernel (B200LP_NO_COOP_PICKS: diagnostic switch back to one
        // k_pick_cluster launch per pick)
        int coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device));
        if (coop && !getenv("B200LP_NO_COOP_PICKS")) {
            const bool ok = cudaFuncSetAttribute(k_blk_picks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)PK_SMEM_MAX) == cudaSuccess &&
                            cudaFuncSetAttribute(k_blk_picks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)PK_SMEM_MAX) == cudaSuccess;
            cudaGetLastError();
            s->coop_picks = ok;
        }
        CKR(s->part_b.ensure(1024));
        CKR(s->pk_err.ensure(1));
        CK(cudaMemsetAsync(s->pk_err.p, 0, sizeof(int32_t), s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));
    *out = s;
    return 0;
}

B200LP_API int b200lp_destroy(b200lp_solver* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (s->graph) cudaGraphExecDestroy(s->graph);
    s->own_T.release();
    s->rowlab.release();
    s->collab.release();
    s->col.release();
    s->part_price.release();
    s->part_ratio.release();
    s->st.release();
    s->h_row.release();
    s->h_col.release();
    s->h_enter.release();
    s->h_leave.release();
    s->sA.release();
    s->sb.release();
    s->sc.release();
    s->sx.release();
    s->sfun.release();
    s->sinfo.release();
    s->sops.release();
    s->sstatus.release();
    s->snpiv.release();
    s->slog.release();
    s->xbuf.release();
    s->gbar.release();
    s->blk_colP.release();
    s->blk_qP.release();
    s->blk_obj.release();
    s->blk_rhs.release();
    s->blk_pend.release();
    if (s->st_host) cudaFreeHost(s->st_host);
    cudaEventDestroy(s->ev0);
    cudaEventDestroy(s->ev1);
    cudaEventDestroy(s->ev_state[0]);
    cudaEventDestroy(s->ev_state[1]);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    delete s;
    return 0;
}

static void drop_graph(b200lp_solver* s) {
    if (s->graph) {
        cudaGraphExecDestroy(s->graph);
        s->graph = nullptr;
        s->graph_key = GraphKey();
    }
}

B200LP_API int b200lp_use_own_stream(b200lp_solver* s) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    s->stream = s->own_stream;
    drop_graph(s);
    return 0;
}

B200LP_API int b200lp_set_stream(b200lp_solver* s, void* stream) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    // the handle is used as given: 0 is the legacy default stream (explicitly requested by the caller; CUDA
    // graphs cannot be captured on it, so the loop falls back to plain launches there)
    s->stream = (cudaStream_t)stream;
    drop_graph(s);
    return 0;
}

B200LP_API int b200lp_synchronize(b200lp_solver* s) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    CKR(set_device(s));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

B200LP_API int b200lp_check_guards(b200lp_solver* s, int64_t* corrupted_bytes) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    if (!corrupted_bytes) return fail(B200LP_E_INVALID, "corrupted_bytes is NULL");
    *corrupted_bytes = 0;
    if (!guard_mode()) return fail(B200LP_E_STATE, "guard mode is off: set B200LP_GUARD=1 before the library is first used");
    CKR(set_device(s));
    CK(cudaDeviceSynchronize());
    long long total = 0;
    bool err = false;
    auto add = [&](long long v) {
        if (v < 0) err = true;
        else total += v;
    };
    add(s->own_T.check()); add(s->rowlab.check()); add(s->collab.check()); add(s->col.check());
    add(s->part_price.check()); add(s->part_ratio.check()); add(s->st.check());
    add(s->h_row.check()); add(s->h_col.check()); add(s->h_enter.check()); add(s->h_leave.check());
    add(s->sA.check()); add(s->sb.check()); add(s->sc.check()); add(s->sx.check()); add(s->sfun.check());
    add(s->sinfo.check()); add(s->sops.check()); add(s->sstatus.check()); add(s->snpiv.check()); add(s->slog.check());
    add(s->snext.check()); add(s->blk_colP.check()); add(s->blk_qP.check()); add(s->blk_obj.check());
    add(s->blk_rhs.check()); add(s->blk_pend.check()); add(s->xbuf.check()); add(s->gbar.check());
    add(s->part_b.check()); add(s->pk_err.check()); add(s->xgen.check()); add(s->shard_ctx.check());
    if (err) return fail(B200LP_E_CUDA, "reading a guard band failed: %s", cudaGetErrorString(cudaGetLastError()));
    *corrupted_bytes = total;
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// tableau binding
// ------------------------------------------------------------------------------------------------------
static int ensure_aux(b200lp_solver* s, int64_t R, int64_t C) {
    CKR(s->rowlab.ensure((size_t)R));
    CKR(s->collab.ensure((size_t)C + 1));
    CKR(s->col.ensure((size_t)R + 64));  // padded: the TMA variant bulk-copies whole row-tile slices
    return 0;
}

static int ensure_hist(b200lp_solver* s, int64_t cap) {
    cap = std::max<int64_t>(cap, 16);
    if (cap > s->hist_cap) {
        CKR(s->h_row.ensure((size_t)cap));
        CKR(s->h_col.ensure((size_t)cap));
        CKR(s->h_enter.ensure((size_t)cap));
        CKR(s->h_leave.ensure((size_t)cap));
        s->hist_cap = cap;
    }
    return 0;
}

static int bind(b200lp_solver* s, double* T, int64_t m, int64_t n_obj, int64_t C, int64_t ld, int64_t n_struct,
                int32_t art_base) {
    if (!T) return fail(B200LP_E_INVALID, "tableau pointer is NULL");
    if (m < 0 || (n_obj != 1 && n_obj != 2) || C < 1) return fail(B200LP_E_INVALID, "bad tableau shape m=%lld n_obj=%lld C=%lld", (long long)m, (long long)n_obj, (long long)C);
    if (ld < C || (ld & 1)) return fail(B200LP_E_INVALID, "row stride ld=%lld must be even and >= C=%lld", (long long)ld, (long long)C);
    if (((uintptr_t)T) & 15) return fail(B200LP_E_INVALID, "tableau base must be 16-byte aligned");
    if (m + n_obj > 0x7fffffff || C > 0x7fffffff) return fail(B200LP_E_INVALID, "tableau dimensions exceed int32 positions");
    s->bind_epoch++;
    s->T = T;
    s->m = m;
    s->n_obj = n_obj;
    s->R = m + n_obj;
    s->C = C;
    s->ld = ld;
    s->n_struct = n_struct;
    s->art_base = art_base;
    CKR(ensure_aux(s, s->R, s->C));
    return 0;
}

B200LP_API int b200lp_attach(b200lp_solver* s, double* T_dev, int64_t m, int64_t n_obj, int64_t C, int64_t ld,
                             int64_t n_struct, int32_t art_base) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    CKR(set_device(s));
    CKR(bind(s, T_dev, m, n_obj, C, ld, n_struct, art_base));
    CK(cudaMemsetAsync(s->st.p, 0, sizeof(DevState), s->stream));
    return 0;
}

B200LP_API int b200lp_dims(b200lp_solver* s, int64_t* m, int64_t* n_obj, int64_t* C, int64_t* ld) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (m) *m = s->m;
    if (n_obj) *n_obj = s->n_obj;
    if (C) *C = s->C;
    if (ld) *ld = s->ld;
    return 0;
}

B200LP_API int b200lp_generate(b200lp_solver* s, uint64_t seed, int64_t n_total, int64_t lab0) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (s->n_obj != 1) return fail(B200LP_E_INVALID, "generated tableaux have one objective row");
    CKR(set_device(s));
    s->n_struct = n_total;
    s->art_base = (int32_t)(n_total + s->m);
    dim3 grid((unsigned)((s->ld / 2 + 255) / 256), (unsigned)std::min<int64_t>(s->R, 2048));
    k_generate<<<grid, 256, 0, s->stream>>>(s->T, s->m, s->R, s->C, s->ld, seed, n_total, lab0, s->rowlab.p, s->collab.p);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(s->st.p, 0, sizeof(DevState), s->stream));
    return 0;
}

B200LP_API int b200lp_set_labels(b200lp_solver* s, const int32_t* rowlab, const int32_t* collab) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    if (rowlab) CK(cudaMemcpyAsync(s->rowlab.p, rowlab, (size_t)s->R * 4, cudaMemcpyHostToDevice, s->stream));
    if (collab) CK(cudaMemcpyAsync(s->collab.p, collab, (size_t)s->C * 4, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

B200LP_API int b200lp_get_labels(b200lp_solver* s, int32_t* rowlab, int32_t* collab) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    if (rowlab) CK(cudaMemcpyAsync(rowlab, s->rowlab.p, (size_t)s->R * 4, cudaMemcpyDeviceToHost, s->stream));
    if (collab) CK(cudaMemcpyAsync(collab, s->collab.p, (size_t)s->C * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------------
static inline int clampi(int64_t v, int64_t lo, int64_t hi) { return (int)std::max(lo, std::min(hi, v)); }

static int launch_flush(b200lp_solver* s) {
    const int blocks = clampi((s->C + PICK_THREADS - 1) / PICK_THREADS, 1, 2 * s->sm_count);
    k_flush_row<<<blocks, PICK_THREADS, 0, s->stream>>>(s->T, s->C, s->ld, s->st.p);
    k_clear_pend<<<1, 1, 0, s->stream>>>(s->st.p);
    s->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

static int launch_price(b200lp_solver* s, int64_t obj_row, int32_t rule, double eps_cost, bool sharded) {
    const int blocks = clampi((s->C + PICK_THREADS * 4 - 1) / (PICK_THREADS * 4), 1, 2 * s->sm_count);
#define PRICE(BL, SH) \
    k_price<BL, SH><<<blocks, PICK_THREADS, 0, s->stream>>>(s->T, s->C, s->ld, obj_row, s->collab.p, s->art_base, eps_cost, s->st.p, s->part_price.p)
    if (rule == B200LP_RULE_BLAND) {
        if (sharded) PRICE(true, true);
        else PRICE(true, false);
    } else {
        if (sharded) PRICE(false, true);
        else PRICE(false, false);
    }
#undef PRICE
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_ratio(b200lp_solver* s, double eps_pivot, bool row_preset, const double* ext, int64_t ext_stride) {
    const int blocks = clampi((s->R + PICK_THREADS - 1) / PICK_THREADS, 1, 2 * s->sm_count);
    if (row_preset)
        k_ratio<true><<<blocks, PICK_THREADS, 0, s->stream>>>(s->T, s->R, s->m, s->C, s->ld, s->rowlab.p, s->collab.p, eps_pivot,
                                                              s->st.p, s->part_ratio.p, s->col.p, ext, ext_stride, s->h_row.p,
                                                              s->h_col.p, s->h_enter.p, s->h_leave.p, s->hist_cap);
    else
        k_ratio<false><<<blocks, PICK_THREADS, 0, s->stream>>>(s->T, s->R, s->m, s->C, s->ld, s->rowlab.p, s->collab.p, eps_pivot,
                                                               s->st.p, s->part_ratio.p, s->col.p, ext, ext_stride, s->h_row.p,
                                                               s->h_col.p, s->h_enter.p, s->h_leave.p, s->hist_cap);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ensure_tmap(b200lp_solver* s, int box_r) {
    if (s->tmap_T == s->T && s->tmap_R == s->R && s->tmap_C == s->C && s->tmap_ld == s->ld && s->tmap_box_r == box_r) return 0;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(B200LP_E_CUDA, "cuTensorMapEncodeTiled not available");
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)s->C, (cuuint64_t)s->R};
    cuuint64_t strides[1] = {(cuuint64_t)s->ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)TMA_BOX_C, (cuuint32_t)box_r};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&s->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)s->T, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200LP_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    s->tmap_T = s->T;
    s->tmap_R = s->R;
    s->tmap_C = s->C;
    s->tmap_ld = s->ld;
    s->tmap_box_r = box_r;
    return 0;
}

// AUTO: the register-resident LDG kernel, except for very wide rows (>= 256 KB), where the TMA pipeline measured
// ~2 % faster on B200 (131072 x 65536: LDG 6.14-6.17 TB/s, TMA 6.28 TB/s; 16384 x 16384: LDG 6.50, TMA 6.30).
static int resolve_variant(const b200lp_solver* s, int32_t v) {
    if (v == B200LP_UPDATE_LDG || v == B200LP_UPDATE_TMA) return v;
    return s->C >= 32768 ? B200LP_UPDATE_TMA : B200LP_UPDATE_LDG;
}

template <int BOX_R, int STAGES, int LAG, int OCC>
static int launch_update_tma(b200lp_solver* s) {
    CKR(ensure_tmap(s, BOX_R));
    const int64_t strips = (s->C + TMA_BOX_C - 1) / TMA_BOX_C;
    const int64_t row_tiles = (s->R + BOX_R - 1) / BOX_R;
    const int64_t ctas = (int64_t)s->sm_count * OCC;
    // ~64 work items per CTA: the static round-robin then loses < 2 % to quantisation (measured best on B200)
    int64_t chunk = std::max<int64_t>(1, row_tiles * strips / (ctas * 64));
    const int64_t n_chunks = (row_tiles + chunk - 1) / chunk;
    const int64_t n_work = strips * n_chunks;
    const int grid = clampi(n_work, 1, ctas);
    k_update_tma<BOX_R, STAGES, LAG, OCC><<<grid, TMA_THREADS, TmaCfg<BOX_R, STAGES>::SMEM_BYTES, s->stream>>>(
        s->tmap, s->T, s->R, s->C, s->ld, s->col.p, s->st.p, (int)strips, (int)chunk, n_work);
    return 0;
}

static int launch_update(b200lp_solver* s, int32_t variant) {
    variant = resolve_variant(s, variant);
    if (variant == B200LP_UPDATE_TMA) {
        CKR((launch_update_tma<TMA_BOX_R, TMA_STAGES, TMA_STORE_LAG, 1>(s)));
    } else {
        const bool wide = s->C >= 2048;
        const int nt = wide ? 256 : 128;
        const int tiles_c = (int)((s->C + 2 * nt - 1) / (2 * nt));
        int64_t tr = s->R * tiles_c / ((int64_t)s->sm_count * 8);
        tr = std::max<int64_t>(8, std::min<int64_t>(64, tr / 8 * 8));
        const int64_t n_tiles = (s->R + tr - 1) / tr * tiles_c;
        const int grid = clampi(n_tiles, 1, (int64_t)s->sm_count * 16);
        if (wide)
            k_update_ldg<256, 8><<<grid, 256, 0, s->stream>>>(s->T, s->R, s->C, s->ld, s->col.p, s->st.p, (int)tr, tiles_c, n_tiles);
        else
            k_update_ldg<128, 8><<<grid, 128, 0, s->stream>>>(s->T, s->R, s->C, s->ld, s->col.p, s->st.p, (int)tr, tiles_c, n_tiles);
    }
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_reset(b200lp_solver* s, int64_t max_pivots, bool keep_count) {
    k_reset_state<<<1, 1, 0, s->stream>>>(s->st.p, (long long)max_pivots, keep_count ? 1 : 0);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

// both decisions of a pivot in one launch on one thread-block cluster (kernels_cluster.cuh)
static int launch_pick_cluster(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, bool blocked) {
    PickArgs A;
    A.T = s->T;
    A.R = s->R;
    A.m = s->m;
    A.C = s->C;
    A.ld = s->ld;
    A.obj_row = obj_row;
    A.rowlab = s->rowlab.p;
    A.collab = s->collab.p;
    A.art_base = s->art_base;
    A.eps_cost = o->eps_cost;
    A.eps_pivot = o->eps_pivot;
    A.st = s->st.p;
    A.col = s->col.p;
    A.B = s->blk;
    A.h_row = s->h_row.p;
    A.h_col = s->h_col.p;
    A.h_enter = s->h_enter.p;
    A.h_leave = s->h_leave.p;
    A.hist_cap = s->hist_cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(s->cluster_ctas);
    cfg.blockDim = dim3(CL_THREADS);
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = s->cluster_ctas;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const bool bland = o->rule == B200LP_RULE_BLAND;
    cudaError_t e;
    if (blocked) e = bland ? cudaLaunchKernelEx(&cfg, k_pick_cluster<true, true>, A) : cudaLaunchKernelEx(&cfg, k_pick_cluster<false, true>, A);
    else e = bland ? cudaLaunchKernelEx(&cfg, k_pick_cluster<true, false>, A) : cudaLaunchKernelEx(&cfg, k_pick_cluster<false, false>, A);
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, "cluster pick launch failed: %s", cudaGetErrorString(e));
    s->launches++;
    return 0;
}

static int launch_snapshot(b200lp_solver* s) {
    if (!s->snaps) return 0;
    const int blocks = clampi((s->R * s->C + 255) / 256, 1, 4 * s->sm_count);
    k_snapshot<<<blocks, 256, 0, s->stream>>>(s->T, s->R, s->C, s->ld, s->st.p, s->snaps, s->snap_cap);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

// one iteration of the loop on the stream
static int enqueue_iteration(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row) {
    if (s->cluster_ctas) {
        CKR(launch_pick_cluster(s, o, obj_row, false));
    } else {
        CKR(launch_price(s, obj_row, o->rule, o->eps_cost, false));
        CKR(launch_ratio(s, o->eps_pivot, false, nullptr, 0));
    }
    CKR(launch_update(s, o->update_variant));
    CKR(launch_snapshot(s));
    return 0;
}

static int enqueue_driveout(b200lp_solver* s, const b200lp_opts* o) {
    CKR(launch_flush(s));
    k_driveout_pick<<<1, 1024, 0, s->stream>>>(s->T, s->m, s->C, s->ld, s->rowlab.p, s->collab.p, s->art_base, o->eps_pivot, s->st.p);
    s->launches++;
    CK(cudaGetLastError());
    CKR(launch_ratio(s, o->eps_pivot, true, nullptr, 0));
    CKR(launch_update(s, o->update_variant));
    CKR(launch_snapshot(s));
    return 0;
}

// ---- look-ahead (blocked) loop ---------------------------------------------------------------------------------
static int blk_setup(b200lp_solver* s) {
    const int64_t Rpad = (s->R + 15) / 16 * 16, Cpad = (s->C + 15) / 16 * 16 + 16;
    CKR(s->blk_colP.ensure((size_t)BLK_KMAX * Rpad));
    CKR(s->blk_qP.ensure((size_t)BLK_KMAX * Cpad));
    CKR(s->blk_obj.ensure((size_t)Cpad));
    CKR(s->blk_rhs.ensure((size_t)Rpad));
    CKR(s->blk_pend.ensure(1));
    s->blk.pend = s->blk_pend.p;
    s->blk.colP = s->blk_colP.p;
    s->blk.qP = s->blk_qP.p;
    s->blk.objcur = s->blk_obj.p;
    s->blk.rhscur = s->blk_rhs.p;
    s->blk.Rpad = Rpad;
    s->blk.Cpad = Cpad;
    return 0;
}

static int blk_block_size(const b200lp_opts* o) {
    return o->check_every > 0 ? (int)std::max(1, std::min(BLK_KMAX, o->check_every)) : BLK_KMAX;
}

static int launch_blk_rowprice(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, bool sharded) {
    const int wb = clampi((std::max(s->R, s->C) + BLK_THREADS * 2 - 1) / (BLK_THREADS * 2), 1, 2 * s->sm_count);
#define ROWPRICE(BL, SH) \
    k_blk_rowprice<BL, SH><<<wb, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->C, s->ld, obj_row, s->collab.p, s->art_base, o->eps_cost, s->st.p, s->part_price.p, s->blk)
    if (o->rule == B200LP_RULE_BLAND) {
        if (sharded) ROWPRICE(true, true);
        else ROWPRICE(true, false);
    } else {
        if (sharded) ROWPRICE(false, true);
        else ROWPRICE(false, false);
    }
#undef ROWPRICE
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_blk_ratio(b200lp_solver* s, const b200lp_opts* o, const double* ext, int64_t ext_stride) {
    const int rb = clampi((s->R + BLK_THREADS - 1) / BLK_THREADS, 1, 2 * s->sm_count);
    k_blk_ratio<<<rb, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->m, s->C, s->ld, s->rowlab.p, s->collab.p, o->eps_pivot, s->st.p,
                                                 s->part_ratio.p, s->blk, ext, ext_stride, s->h_row.p, s->h_col.p,
                                                 s->h_enter.p, s->h_leave.p, s->hist_cap);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

// one look-ahead pivot: one launch on a thread-block cluster, or [row part of the previous pivot + pricing] and
// [ratio on the replayed column]
static int enqueue_blk_pick(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row) {
    if (s->cluster_ctas) return launch_pick_cluster(s, o, obj_row, true);
    CKR(launch_blk_rowprice(s, o, obj_row, false));
    CKR(launch_blk_ratio(s, o, nullptr, 0));
    return 0;
}

// Can the K picks of a block run as ONE cooperative kernel?  Every CTA keeps the history of its C / G columns and
// R / G rows in shared memory, and all G = #SMs CTAs must be co-resident.
static bool coop_picks_plan(const b200lp_solver* s, int K, int* wC, int* wR, size_t* smem) {
    if (!s->coop_picks || s->sm_count > 1024) return false;
    const int G = s->sm_count;
    *wC = (int)((s->C + G - 1) / G);
    *wR = (int)((s->R + G - 1) / G);
    *smem = (size_t)K * (*wC + *wR) * 8 + (size_t)(*wC + *wR) * 4 + 16;
    return *smem <= PK_SMEM_MAX;
}

static int launch_blk_picks(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int K, int wC, int wR, size_t smem) {
    PicksArgs P;
    P.A.T = s->T;
    P.A.R = s->R;
    P.A.m = s->m;
    P.A.C = s->C;
    P.A.ld = s->ld;
    P.A.obj_row = obj_row;
    P.A.rowlab = s->rowlab.p;
    P.A.collab = s->collab.p;
    P.A.art_base = s->art_base;
    P.A.eps_cost = o->eps_cost;
    P.A.eps_pivot = o->eps_pivot;
    P.A.st = s->st.p;
    P.A.col = s->col.p;
    P.A.B = s->blk;
    P.A.h_row = s->h_row.p;
    P.A.h_col = s->h_col.p;
    P.A.h_enter = s->h_enter.p;
    P.A.h_leave = s->h_leave.p;
    P.A.hist_cap = s->hist_cap;
    P.K = K;
    P.wC = wC;
    P.wR = wR;
    P.barrier = s->gbar.p;
    P.partA = s->part_price.p;
    P.partB = s->part_b.p;
    P.error = s->pk_err.p;
    CK(cudaMemsetAsync(s->gbar.p, 0, sizeof(unsigned long long), s->stream));
    void* args[] = {&P};
    const void* fn = o->rule == B200LP_RULE_BLAND ? (const void*)k_blk_picks<true> : (const void*)k_blk_picks<false>;
    CK(cudaLaunchCooperativeKernel(fn, dim3(s->sm_count), dim3(PK_THREADS), args, smem, s->stream));
    s->launches++;
    return 0;
}

static int enqueue_blk_flush(b200lp_solver* s, int K, int64_t obj_row, bool row_done = false) {
    // the row part of the block's last pivot (the fused kernel would only do it at the next pick)
    if (!row_done) {
        const int wb = clampi((std::max(s->R, s->C) + BLK_THREADS - 1) / BLK_THREADS, 1, 2 * s->sm_count);
        k_blk_row<<<wb, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->C, s->ld, obj_row, s->st.p, s->blk);
        s->launches++;
    }
    // pivot rows / pivot column pairs first (general replay), then everything else (plain FMA chains)
    const unsigned sp_blocks = (unsigned)((std::max(s->R, s->C) + 255) / 256);
    k_blk_flush_special<<<dim3(sp_blocks, 2), 256, 0, s->stream>>>(s->T, s->R, s->C, s->ld, s->st.p, s->blk);
    if (!s->flush_mma) {
        const int64_t sw = 64 * FL_NP * FL_WC, n_strips = (s->C + sw - 1) / sw, n_rb = (s->R + FL_TR - 1) / FL_TR;
        const int grid = clampi(n_strips * n_rb, 1, s->sm_count);
        k_blk_flush_db<FL_WC, FL_TR, FL_NP><<<grid, 256, FL_SMEM_BYTES, s->stream>>>(s->T, s->R, s->C, s->ld, s->st.p,
                                                                                     s->blk, n_rb, n_strips);
    } else {  // four pending steps of an 8 x 8 tile per DMMA instruction (kernels_flush_mma.cuh)
        const int64_t n_strips = (s->C + FM_SW - 1) / FM_SW, n_rb = (s->R + FM_TR - 1) / FM_TR;
        const int grid = clampi(n_strips * n_rb, 1, s->sm_count);
        k_blk_flush_mma<<<grid, 256, FM_SMEM_BYTES, s->stream>>>(s->T, s->R, s->C, s->ld, s->st.p, s->blk, n_rb, n_strips);
    }
    k_blk_clear<<<1, 1, 0, s->stream>>>(s->st.p, s->blk);
    s->launches += 3;
    (void)K;
    CK(cudaGetLastError());
    return 0;
}

static int enqueue_blk_block(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int K) {
    int wC, wR;
    size_t smem;
    if (coop_picks_plan(s, K, &wC, &wR, &smem)) {
        CKR(launch_blk_picks(s, o, obj_row, K, wC, wR, smem));
        CKR(enqueue_blk_flush(s, K, obj_row, true));
        return 0;
    }
    for (int k = 0; k < K; ++k) CKR(enqueue_blk_pick(s, o, obj_row));
    CKR(enqueue_blk_flush(s, K, obj_row));
    return 0;
}

static int default_check_every(const b200lp_solver* s) {
    // enough work per replay to hide the host's status read; big tableaux need few pivots per replay
    const double bytes = 16.0 * (double)s->R * (double)s->C;
    if (bytes > 1e9) return 8;
    if (bytes > 5e7) return 32;
    return 64;
}

static int get_graph(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int iters, int blocked_k = 0) {
    GraphKey k;
    k.blocked_k = blocked_k;
    k.T = s->T;
    k.R = s->R;
    k.C = s->C;
    k.ld = s->ld;
    k.obj_row = obj_row;
    k.rule = o->rule;
    k.variant = resolve_variant(s, o->update_variant);
    k.iters = iters;
    k.eps_cost = o->eps_cost;
    k.eps_pivot = o->eps_pivot;
    k.hist_cap = s->hist_cap;
    k.snaps = s->snaps;
    k.snap_cap = s->snap_cap;
    k.m = s->m;
    k.art_base = s->art_base;
    const void* baked[12] = {s->rowlab.p, s->collab.p, s->col.p, s->h_row.p, s->h_col.p, s->h_enter.p, s->h_leave.p,
                             s->blk.pend, s->blk.colP, s->blk.qP, s->blk.objcur, s->blk.rhscur};
    memcpy(k.ptrs, baked, sizeof(baked));
    if (s->graph && k == s->graph_key) return 0;
    if (s->graph) {
        cudaGraphExecDestroy(s->graph);
        s->graph = nullptr;
    }
    const int64_t before = s->launches;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    for (int i = 0; i < iters && !rc; ++i)
        rc = blocked_k ? enqueue_blk_block(s, o, obj_row, blocked_k) : enqueue_iteration(s, o, obj_row);
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    s->launches = before;  // capture launched nothing
    if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    s->graph_key = k;
    return 0;
}

static int read_state(b200lp_solver* s, DevState* out);

struct OnchipPlan {
    int G = 0, stride = 0;
    size_t smem = 0;
};

static bool onchip_plan(const b200lp_solver* s, OnchipPlan* plan) {
    if (s->C < 2 || s->snaps) return false;
    // one CTA per SM (measured on B200: 32..148 CTAs cost the same fixed latency per pivot; more CTAs = smaller slices)
    const int64_t G = std::min<int64_t>(s->sm_count, s->C - 1);
    const int64_t wmax = (s->C - 1 + G - 1) / G;
    int64_t stride = wmax + 1;
    if ((stride & 1) == 0) ++stride;
    const size_t smem = (size_t)(s->R * stride + s->R + stride) * 8 + (size_t)(s->R + wmax) * 4 + 64;
    if (smem > ONCHIP_SMEM_MAX) return false;
    plan->G = (int)G;
    plan->stride = (int)stride;
    plan->smem = smem;
    return true;
}

// One phase of the loop as ONE persistent cooperative kernel (kernels_onchip.cuh).
static int run_onchip(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, const OnchipPlan& plan, DevState* final_state) {
    CKR(launch_flush(s));
    const size_t xdoubles = (size_t)2 * plan.G * onchip_xstride(s->R);
    CKR(s->xbuf.ensure(xdoubles));
    CK(cudaMemsetAsync(s->gbar.p, 0, sizeof(unsigned long long), s->stream));
    OnchipParams P;
    P.T = s->T;
    P.R = s->R;
    P.m = s->m;
    P.C = s->C;
    P.ld = s->ld;
    P.obj_row = obj_row;
    P.rowlab = s->rowlab.p;
    P.collab = s->collab.p;
    P.art_base = s->art_base;
    P.rule = o->rule;
    P.eps_cost = o->eps_cost;
    P.eps_pivot = o->eps_pivot;
    P.st = s->st.p;
    P.xbuf = s->xbuf.p;
    P.barrier = s->gbar.p;
    P.h_row = s->h_row.p;
    P.h_col = s->h_col.p;
    P.h_enter = s->h_enter.p;
    P.h_leave = s->h_leave.p;
    P.hist_cap = s->hist_cap;
    P.stride = plan.stride;
    P.error = s->pk_err.p;
    P.time_budget_ns = 0;  // the kernel runs a whole phase: it checks the wall-clock bound itself (%globaltimer)
    if (s->deadline > 0.0) P.time_budget_ns = (long long)std::max(1.0, (s->deadline - now_s()) * 1e9);
    void* args[] = {&P};
    CK(cudaLaunchCooperativeKernel((void*)k_solve_onchip, dim3(plan.G), dim3(ONCHIP_THREADS), args, plan.smem, s->stream));
    s->launches++;
    CK(cudaMemcpyAsync(&s->st_host[0], s->st.p, sizeof(DevState), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *final_state = s->st_host[0];
    if (final_state->status == B200LP_STATUS_NUMERICAL) {
        int32_t err = 0;
        CK(cudaMemcpy(&err, s->pk_err.p, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CK(cudaMemset(s->pk_err.p, 0, sizeof(err)));
            return fail(B200LP_E_CUDA, "on-chip loop: a grid barrier timed out");
        }
    }
    return 0;
}

// Runs chunks of iterations until the device reports done.  mode 0: pricing loop on obj_row; mode 1: drive-out.
static int run_loop(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int mode, DevState* final_state) {
    if (deadline_passed(s)) {  // nothing is enqueued once the wall-clock bound of the call has expired
        CKR(read_state(s, final_state));
        final_state->done = 1;
        final_state->status = B200LP_STATUS_LIMIT;
        return 0;
    }
    if (mode == 0 && o->loop_mode == B200LP_LOOP_AUTO) {
        OnchipPlan plan;
        if (onchip_plan(s, &plan)) return run_onchip(s, o, obj_row, plan, final_state);
    }
    // AUTO: whatever does not fit the on-chip loop takes the look-ahead loop -- bit-identical pivots, one tableau pass
    // per K pivots.  Measured (scripts/probe_sizes.py, Bland, pivots/s, rank-1 graph loop vs look-ahead K = 32):
    // 2048^2 53k vs 92k, 4096^2 19k vs 83k, 8192^2 5.7k vs 58k, 16384^2 1.5k vs 25k (on-chip at 1536^2: 120k vs 91k).
    const bool want_blocked = o->loop_mode == B200LP_LOOP_BLOCKED || o->loop_mode == B200LP_LOOP_AUTO;
    const int blocked_k = (mode == 0 && want_blocked && !s->snaps) ? blk_block_size(o) : 0;
    int iters = o->check_every > 0 ? o->check_every : default_check_every(s);
    if (mode == 1) iters = std::min(iters, 8);
    if (blocked_k) {
        iters = 4;  // look-ahead blocks per replay
        CKR(launch_flush(s));
        CKR(blk_setup(s));
        const int ib = clampi((std::max(s->R, s->C) + BLK_THREADS - 1) / BLK_THREADS, 1, 2 * s->sm_count);
        k_blk_init<<<ib, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->C, s->ld, obj_row, s->st.p, s->blk);
        s->launches++;
        CK(cudaGetLastError());
    }
    bool use_graph = o->loop_mode != B200LP_LOOP_LAUNCHES && mode == 0 && s->stream != (cudaStream_t)0;
    int wC = 0, wR = 0;
    size_t pk_smem = 0;
    const bool coop = blocked_k && coop_picks_plan(s, blocked_k, &wC, &wR, &pk_smem);
    if (use_graph) {
        const int rc = get_graph(s, o, obj_row, iters, blocked_k);
        if (rc && coop) {  // a driver that cannot capture a cooperative launch: 4 launches per block need no graph
            cudaGetLastError();
            use_graph = false;
        } else if (rc) {
            return rc;
        }
    }
    const int picks = s->cluster_ctas ? 1 : 2;  // launches per pick
    const int per_iter = blocked_k ? (coop ? 4 : picks * blocked_k + 4) : picks + 1 + (s->snaps ? 1 : 0);
    int slot = 0;
    bool first = true, timed_out = false;
    for (;;) {
        if (use_graph) {
            CK(cudaGraphLaunch(s->graph, s->stream));
            s->launches += (int64_t)per_iter * iters;
        } else {
            for (int i = 0; i < iters; ++i) {
                if (blocked_k) CKR(enqueue_blk_block(s, o, obj_row, blocked_k));
                else if (mode == 0) CKR(enqueue_iteration(s, o, obj_row));
                else CKR(enqueue_driveout(s, o));
            }
        }
        CK(cudaMemcpyAsync(&s->st_host[slot], s->st.p, sizeof(DevState), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaEventRecord(s->ev_state[slot], s->stream));
        if (!first) {
            // look at the PREVIOUS chunk's state while this chunk runs
            CK(cudaEventSynchronize(s->ev_state[slot ^ 1]));
            if (s->st_host[slot ^ 1].done) break;
        }
        first = false;
        if (deadline_passed(s)) {  // whole replays only: the tableau is consistent where the loop stops
            timed_out = true;
            break;
        }
        slot ^= 1;
    }
    CK(cudaStreamSynchronize(s->stream));
    // the newest copy is at least as recent as the one that reported done
    const DevState& a = s->st_host[slot];
    *final_state = (a.done || timed_out) ? a : s->st_host[slot ^ 1];
    if (timed_out && !final_state->done) {
        final_state->done = 1;
        final_state->status = B200LP_STATUS_LIMIT;
    }
    if (coop && final_state->status == B200LP_STATUS_NUMERICAL) {
        int32_t err = 0;
        CK(cudaMemcpy(&err, s->pk_err.p, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CK(cudaMemset(s->pk_err.p, 0, sizeof(err)));
            return fail(B200LP_E_CUDA, err == 1 ? "look-ahead pick kernel: a grid barrier timed out"
                                                : "look-ahead pick kernel launched with pivots pending");
        }
    }
    return 0;
}

// run_loop for a pricing phase, with the Dantzig -> Bland continuation described above.  `o` is the call's private
// copy of the options (its rule may be switched), `budget` the current cap held in DevState.max_pivots.
static int run_phase_fb(b200lp_solver* s, b200lp_opts* o, int64_t obj_row, bool is_auto, int64_t cap, int64_t* budget,
                        DevState* fin) {
    CKR(run_loop(s, o, obj_row, 0, fin));
    if (fin->status == B200LP_STATUS_LIMIT && is_auto && o->rule == B200LP_RULE_DANTZIG && !deadline_passed(s)) {
        o->rule = B200LP_RULE_BLAND;
        *budget += cap;
        CKR(launch_reset(s, *budget, true));
        CKR(run_loop(s, o, obj_row, 0, fin));
    }
    return 0;
}

static int copy_history(b200lp_solver* s, b200lp_result* r, int64_t n) {
    if (!r || r->hist_cap <= 0) return 0;
    const int64_t k = std::min<int64_t>(std::min<int64_t>(n, r->hist_cap), s->hist_cap);
    if (k <= 0) return 0;
    if (r->piv_row) CK(cudaMemcpyAsync(r->piv_row, s->h_row.p, (size_t)k * 4, cudaMemcpyDeviceToHost, s->stream));
    if (r->piv_col) CK(cudaMemcpyAsync(r->piv_col, s->h_col.p, (size_t)k * 4, cudaMemcpyDeviceToHost, s->stream));
    if (r->enter_lab) CK(cudaMemcpyAsync(r->enter_lab, s->h_enter.p, (size_t)k * 4, cudaMemcpyDeviceToHost, s->stream));
    if (r->leave_lab) CK(cudaMemcpyAsync(r->leave_lab, s->h_leave.p, (size_t)k * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

static int read_solution_impl(b200lp_solver* s, double* x_host, double* fun) {
    CKR(launch_flush(s));
    CKR(s->sx.ensure((size_t)std::max<int64_t>(1, s->n_struct)));
    CK(cudaMemsetAsync(s->sx.p, 0, (size_t)std::max<int64_t>(1, s->n_struct) * 8, s->stream));
    const int blocks = clampi((std::max<int64_t>(s->m, 1) + 255) / 256, 1, 1 << 30);
    k_read_solution<<<blocks, 256, 0, s->stream>>>(s->T, s->m, s->C, s->ld, s->rowlab.p, s->n_struct, s->sx.p, s->sfun.p);
    s->launches++;
    CK(cudaGetLastError());
    if (x_host && s->n_struct > 0) CK(cudaMemcpyAsync(x_host, s->sx.p, (size_t)s->n_struct * 8, cudaMemcpyDeviceToHost, s->stream));
    if (fun) CK(cudaMemcpyAsync(fun, s->sfun.p, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

B200LP_API int b200lp_read_solution(b200lp_solver* s, double* x_host, double* fun) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    return read_solution_impl(s, x_host, fun);
}

B200LP_API int b200lp_read_tableau(b200lp_solver* s, double* T_host) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!T_host) return fail(B200LP_E_INVALID, "T_host is NULL");
    CKR(set_device(s));
    CKR(launch_flush(s));
    CK(cudaMemcpy2DAsync(T_host, (size_t)s->C * 8, s->T, (size_t)s->ld * 8, (size_t)s->C * 8, (size_t)s->R,
                         cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

// Pivot budget (same rule as oracle/simplex_oracle.c).  max_pivots >= 2^40 means "automatic": 200 * (m + C) + 10000
// pivots, and -- because Dantzig's rule with lowest-id tie-breaking can cycle on degenerate problems -- a phase that
// exhausts an AUTOMATIC budget under Dantzig continues from the current basis under Bland's rule (which cannot cycle)
// with one more budget.  A device loop must always terminate; the reference bounds its solver by wall clock instead
// (time_limit = 10 s, solver_controller.py:76).  An explicit budget is honoured as given (status LIMIT).
static const int64_t AUTO_BUDGET = (int64_t)1 << 40;
static int64_t auto_cap(int64_t m, int64_t C) { return 200 * (m + C) + 10000; }

static int check_opts(const b200lp_opts* o) {
    if (!o) return fail(B200LP_E_INVALID, "opts is NULL");
    if (o->rule != B200LP_RULE_DANTZIG && o->rule != B200LP_RULE_BLAND) return fail(B200LP_E_INVALID, "unknown rule %d", o->rule);
    if (o->update_variant < 0 || o->update_variant > 2) return fail(B200LP_E_INVALID, "unknown update variant %d", o->update_variant);
    if (o->max_pivots < 0) return fail(B200LP_E_INVALID, "max_pivots < 0");
    if (o->loop_mode < 0 || o->loop_mode > 3) return fail(B200LP_E_INVALID, "unknown loop_mode %d", o->loop_mode);
    if (o->time_limit_s != o->time_limit_s) return fail(B200LP_E_INVALID, "time_limit_s is NaN");
    return 0;
}

B200LP_API int b200lp_run(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, b200lp_result* r) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(check_opts(o));
    if (obj_row < s->m || obj_row >= s->R) return fail(B200LP_E_INVALID, "obj_row %lld is not an objective row", (long long)obj_row);
    CKR(set_device(s));
    arm_deadline(s, o);
    const int64_t l0 = s->launches;
    b200lp_opts oo = *o;
    const bool is_auto = o->max_pivots >= AUTO_BUDGET;
    const int64_t cap = is_auto ? auto_cap(s->m, s->C) : o->max_pivots;
    int64_t budget = cap;
    CKR(ensure_hist(s, std::min<int64_t>(is_auto ? 2 * cap : cap, r && r->hist_cap > 0 ? r->hist_cap : 16)));
    CKR(launch_reset(s, budget, false));
    DevState fin;
    CK(cudaEventRecord(s->ev0, s->stream));
    CKR(run_phase_fb(s, &oo, obj_row, is_auto, cap, &budget, &fin));
    CKR(launch_flush(s));
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (r) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        r->status = fin.status;
        r->n_pivots = fin.n_pivots;
        r->n_phase1 = 0;
        r->device_ms = ms;
        CKR(read_solution_impl(s, r->x_len >= s->n_struct ? r->x : nullptr, &r->fun));
        CKR(copy_history(s, r, fin.n_pivots));
        r->kernel_launches = s->launches - l0;
    }
    return 0;
}

B200LP_API int b200lp_solve(b200lp_solver* s, const b200lp_opts* o, b200lp_result* r) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(check_opts(o));
    CKR(set_device(s));
    arm_deadline(s, o);
    const int64_t l0 = s->launches;
    b200lp_opts oo = *o;
    const bool is_auto = o->max_pivots >= AUTO_BUDGET;
    const int64_t cap = is_auto ? auto_cap(s->m, s->C) : o->max_pivots;
    int64_t budget = cap;
    CKR(ensure_hist(s, std::min<int64_t>(is_auto ? 2 * cap : cap, r && r->hist_cap > 0 ? r->hist_cap : 16)));
    CKR(launch_reset(s, budget, false));
    DevState fin;
    memset(&fin, 0, sizeof(fin));
    int status = B200LP_STATUS_OPTIMAL;
    int64_t n_phase1 = 0;
    CK(cudaEventRecord(s->ev0, s->stream));
    if (s->n_obj == 2) {
        CKR(run_phase_fb(s, &oo, s->m + 1, is_auto, cap, &budget, &fin));
        status = fin.status;
        if (status == B200LP_STATUS_UNBOUNDED) status = B200LP_STATUS_NUMERICAL;
        if (status == B200LP_STATUS_OPTIMAL) {
            CKR(launch_flush(s));
            double w = 0.0;
            CK(cudaMemcpyAsync(&w, s->T + (s->m + 1) * s->ld + s->C - 1, 8, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            if (w < -o->eps_feas) status = B200LP_STATUS_INFEASIBLE;
        }
        if (status == B200LP_STATUS_OPTIMAL) {
            CKR(launch_reset(s, budget, true));
            CKR(run_loop(s, &oo, s->m, 1, &fin));
            status = fin.status;
        }
        n_phase1 = fin.n_pivots;
    }
    if (status == B200LP_STATUS_OPTIMAL) {
        CKR(launch_reset(s, budget, true));
        CKR(run_phase_fb(s, &oo, s->m, is_auto, cap, &budget, &fin));
        status = fin.status;
    }
    CKR(launch_flush(s));
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (r) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        r->status = status;
        r->n_pivots = fin.n_pivots;
        r->n_phase1 = n_phase1;
        r->device_ms = ms;
        CKR(read_solution_impl(s, r->x_len >= s->n_struct ? r->x : nullptr, &r->fun));
        CKR(copy_history(s, r, fin.n_pivots));
        r->kernel_launches = s->launches - l0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// one LP from arrays (the linprog seam)
// ------------------------------------------------------------------------------------------------------
B200LP_API int b200lp_build_dense(b200lp_solver* s, const b200lp_problem* p) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    if (!p) return fail(B200LP_E_INVALID, "problem is NULL");
    const int64_t m = p->m, n = p->n;
    if (m < 0 || n < 0) return fail(B200LP_E_INVALID, "negative dimensions");
    if (n > 0 && !p->c) return fail(B200LP_E_INVALID, "c is NULL");
    if (m > 0 && (!p->b || !p->ops || (n > 0 && !p->A))) return fail(B200LP_E_INVALID, "A, b or ops is NULL");
    const int64_t lda = p->lda > 0 ? p->lda : n;
    if (lda < n) return fail(B200LP_E_INVALID, "lda < n");
    CKR(set_device(s));
    const int64_t l0 = s->launches;

    // b decides the row normalisation; with device inputs it is the only array read back (m doubles)
    std::vector<double> bh((size_t)m);
    if (m > 0) {
        if (p->on_device) {
            CK(cudaMemcpyAsync(bh.data(), p->b, (size_t)m * 8, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
        } else {
            memcpy(bh.data(), p->b, (size_t)m * 8);
        }
    }
    std::vector<RowInfo> info((size_t)m);
    std::vector<int32_t> rowlab((size_t)m + 2, -1);
    int64_t n_ge = 0, n_art = 0;
    const int32_t art_base = (int32_t)(n + m);
    for (int64_t i = 0; i < m; ++i) {
        int op = p->ops[i];
        if (op < 0 || op > 2) return fail(B200LP_E_INVALID, "ops[%lld] = %d is not L/G/E", (long long)i, op);
        const bool neg = bh[(size_t)i] < 0.0;
        if (neg && op != B200LP_OP_EQ) op = (op == B200LP_OP_LE) ? B200LP_OP_GE : B200LP_OP_LE;
        info[(size_t)i].flags = (neg ? 1 : 0) | (op << 1);
        info[(size_t)i].surplus = -1;
        if (op == B200LP_OP_GE) info[(size_t)i].surplus = (int32_t)(n + n_ge++);
        if (op != B200LP_OP_LE) ++n_art;
        rowlab[(size_t)i] = (op == B200LP_OP_LE) ? (int32_t)(n + i) : (int32_t)(art_base + i);
    }
    const int64_t n_obj = n_art > 0 ? 2 : 1;
    const int64_t C = n + n_ge + 1;
    const int64_t R = m + n_obj;
    // row stride: a multiple of 2 KB for tableaux that stream from HBM (a stride of 16 doubles' granularity costs the update
    // kernels 7-8 % of the bandwidth on B200, scripts/probe_shard_shapes.py), 128 bytes for the small ones
    const int64_t ld = C >= 4096 ? (C + 255) / 256 * 256 : (C + 15) / 16 * 16;
    std::vector<int32_t> collab((size_t)C, -1);
    for (int64_t j = 0; j < n; ++j) collab[(size_t)j] = (int32_t)j;
    for (int64_t i = 0; i < m; ++i)
        if (info[(size_t)i].surplus >= 0) collab[(size_t)info[(size_t)i].surplus] = (int32_t)(n + i);

    CKR(s->own_T.ensure((size_t)(R * ld)));
    CKR(bind(s, s->own_T.p, m, n_obj, C, ld, n, art_base));
    CKR(s->sinfo.ensure((size_t)std::max<int64_t>(m, 1)));
    const double *dA = p->A, *db = p->b, *dc = p->c;
    int64_t dlda = lda;
    if (!p->on_device) {
        CKR(s->sA.ensure((size_t)std::max<int64_t>(m * n, 1)));
        CKR(s->sb.ensure((size_t)std::max<int64_t>(m, 1)));
        CKR(s->sc.ensure((size_t)std::max<int64_t>(n, 1)));
        if (m > 0 && n > 0)
            CK(cudaMemcpy2DAsync(s->sA.p, (size_t)n * 8, p->A, (size_t)lda * 8, (size_t)n * 8, (size_t)m, cudaMemcpyHostToDevice, s->stream));
        if (m > 0) CK(cudaMemcpyAsync(s->sb.p, p->b, (size_t)m * 8, cudaMemcpyHostToDevice, s->stream));
        if (n > 0) CK(cudaMemcpyAsync(s->sc.p, p->c, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
        dA = s->sA.p;
        db = s->sb.p;
        dc = s->sc.p;
        dlda = n;
    }
    if (m > 0) CK(cudaMemcpyAsync(s->sinfo.p, info.data(), (size_t)m * sizeof(RowInfo), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->rowlab.p, rowlab.data(), (size_t)R * 4, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->collab.p, collab.data(), (size_t)C * 4, cudaMemcpyHostToDevice, s->stream));
    {
        dim3 grid((unsigned)((ld + 255) / 256), (unsigned)std::min<int64_t>(m + 1, 4096));
        k_build_rows<<<grid, 256, 0, s->stream>>>(s->T, m, n, C, ld, dA, dlda, db, dc, s->sinfo.p);
        s->launches++;
        CK(cudaGetLastError());
        if (n_obj == 2) {
            k_build_phase1_row<<<(unsigned)((ld + 255) / 256), 256, 0, s->stream>>>(s->T, m, C, ld, s->rowlab.p, art_base);
            s->launches++;
            CK(cudaGetLastError());
        }
    }
    // the host vectors above must outlive the async copies
    CK(cudaStreamSynchronize(s->stream));
    (void)l0;
    return 0;
}

B200LP_API int b200lp_solve_dense(b200lp_solver* s, const b200lp_problem* p, const b200lp_opts* o, b200lp_result* r) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    if (!p || !r) return fail(B200LP_E_INVALID, "problem/result is NULL");
    CKR(check_opts(o));
    const int64_t l0 = s->launches;
    s->deadline_outer = o->time_limit_s > 0.0 ? now_s() + o->time_limit_s : 0.0;  // the copies and the build count
    int rc = b200lp_build_dense(s, p);
    const int64_t built = s->launches - l0;
    if (!rc) rc = b200lp_solve(s, o, r);
    s->deadline_outer = 0.0;
    if (rc) return rc;
    r->kernel_launches += built;
    return 0;
}

B200LP_API int b200lp_set_snapshots(b200lp_solver* s, double* snaps_dev, int64_t cap) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    s->snaps = (snaps_dev && cap > 0) ? snaps_dev : nullptr;
    s->snap_cap = s->snaps ? cap : 0;
    s->bind_epoch++;
    drop_graph(s);  // the snapshot buffer is baked into the captured launches (and allocators hand addresses out again)
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// single-phase entry points
// ------------------------------------------------------------------------------------------------------
static int read_state(b200lp_solver* s, DevState* out) {
    CK(cudaMemcpyAsync(&s->st_host[0], s->st.p, sizeof(DevState), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *out = s->st_host[0];
    return 0;
}

B200LP_API int b200lp_select_entering(b200lp_solver* s, int64_t obj_row, int32_t rule, double eps_cost, int64_t* col_out) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (obj_row < s->m || obj_row >= s->R) return fail(B200LP_E_INVALID, "obj_row %lld is not an objective row", (long long)obj_row);
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(launch_reset(s, (int64_t)1 << 40, true));
    CKR(launch_price(s, obj_row, rule, eps_cost, false));
    DevState st;
    CKR(read_state(s, &st));
    if (col_out) *col_out = st.have_pivot ? st.s : -1;
    return 0;
}

B200LP_API int b200lp_ratio_test(b200lp_solver* s, int64_t col, double eps_pivot, int64_t* row_out) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (col < 0 || col >= s->C - 1) return fail(B200LP_E_INVALID, "column %lld out of range", (long long)col);
    CKR(set_device(s));
    CKR(launch_flush(s));
    // dry run: take the decision on a scratch copy of the labels so that nothing is swapped or logged
    std::vector<int32_t> rl((size_t)s->R), cl((size_t)s->C);
    CKR(b200lp_get_labels(s, rl.data(), cl.data()));
    DevState before;
    CKR(read_state(s, &before));
    CKR(ensure_hist(s, 16));
    k_set_pivot<<<1, 1, 0, s->stream>>>(s->st.p, -1, (int32_t)col, s->collab.p);
    s->launches++;
    CKR(launch_ratio(s, eps_pivot, false, nullptr, 0));
    DevState st;
    CKR(read_state(s, &st));
    if (row_out) *row_out = (st.status == B200LP_STATUS_UNBOUNDED && st.done) ? -1 : st.r;
    // undo the bookkeeping of k_ratio
    CKR(b200lp_set_labels(s, rl.data(), cl.data()));
    before.ticket_price = before.ticket_ratio = 0;
    CK(cudaMemcpyAsync(s->st.p, &before, sizeof(DevState), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

B200LP_API int b200lp_pivot(b200lp_solver* s, int64_t row, int64_t col, int32_t update_variant) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (row < 0 || row >= s->m) return fail(B200LP_E_INVALID, "row %lld out of range", (long long)row);
    if (col < 0 || col >= s->C - 1) return fail(B200LP_E_INVALID, "column %lld out of range", (long long)col);
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(ensure_hist(s, 16));
    k_set_pivot<<<1, 1, 0, s->stream>>>(s->st.p, (int32_t)row, (int32_t)col, s->collab.p);
    s->launches++;
    CKR(launch_ratio(s, 0.0, true, nullptr, 0));
    CKR(launch_update(s, update_variant));
    CKR(launch_flush(s));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

B200LP_API int b200lp_time_update(b200lp_solver* s, int64_t row, int64_t col, int32_t update_variant, int32_t reps,
                                  double* ms_per_launch) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (row < 0 || row >= s->m || col < 0 || col >= s->C - 1 || reps < 1) return fail(B200LP_E_INVALID, "bad arguments");
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(ensure_hist(s, 16));
    k_set_pivot<<<1, 1, 0, s->stream>>>(s->st.p, (int32_t)row, (int32_t)col, s->collab.p);
    s->launches++;
    CKR(launch_ratio(s, 0.0, true, nullptr, 0));
    CKR(launch_update(s, update_variant));  // warm-up
    CK(cudaEventRecord(s->ev0, s->stream));
    for (int i = 0; i < reps; ++i) CKR(launch_update(s, update_variant));
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    if (ms_per_launch) *ms_per_launch = (double)ms / reps;
    CKR(launch_flush(s));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

// `iters` iterations of the loop with plain launches and CUDA events around every kernel: the live per-kernel
// durations that bench.py reports in "roofline" (kernels in their real order, caches in their real state).
B200LP_API int b200lp_profile_loop(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int32_t iters,
                                   double* ms_price, double* ms_ratio, double* ms_update, int64_t* pivots_done) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(check_opts(o));
    if (iters < 1 || iters > 4096) return fail(B200LP_E_INVALID, "iters out of range");
    if (obj_row < s->m || obj_row >= s->R) return fail(B200LP_E_INVALID, "obj_row %lld is not an objective row", (long long)obj_row);
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(ensure_hist(s, 16));
    CKR(launch_reset(s, iters, false));
    std::vector<cudaEvent_t> ev((size_t)iters * 4);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(ev[(size_t)i * 4 + 0], s->stream));
        if (s->cluster_ctas) {  // one launch takes both decisions: reported as "price", "ratio" = 0
            CKR(launch_pick_cluster(s, o, obj_row, false));
            CK(cudaEventRecord(ev[(size_t)i * 4 + 1], s->stream));
        } else {
            CKR(launch_price(s, obj_row, o->rule, o->eps_cost, false));
            CK(cudaEventRecord(ev[(size_t)i * 4 + 1], s->stream));
            CKR(launch_ratio(s, o->eps_pivot, false, nullptr, 0));
        }
        CK(cudaEventRecord(ev[(size_t)i * 4 + 2], s->stream));
        CKR(launch_update(s, o->update_variant));
        CK(cudaEventRecord(ev[(size_t)i * 4 + 3], s->stream));
    }
    CKR(launch_flush(s));
    DevState st;
    CKR(read_state(s, &st));
    double tp = 0, tr = 0, tu = 0;
    const int64_t done = std::max<int64_t>(1, std::min<int64_t>(st.n_pivots, iters));
    for (int64_t i = 0; i < done; ++i) {
        float a = 0, b = 0, c = 0;
        CK(cudaEventElapsedTime(&a, ev[(size_t)i * 4 + 0], ev[(size_t)i * 4 + 1]));
        CK(cudaEventElapsedTime(&b, ev[(size_t)i * 4 + 1], ev[(size_t)i * 4 + 2]));
        CK(cudaEventElapsedTime(&c, ev[(size_t)i * 4 + 2], ev[(size_t)i * 4 + 3]));
        tp += a;
        tr += b;
        tu += c;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (ms_price) *ms_price = tp / done;
    if (ms_ratio) *ms_ratio = tr / done;
    if (ms_update) *ms_update = tu / done;
    if (pivots_done) *pivots_done = st.n_pivots;
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// column-sharded tableau
// ------------------------------------------------------------------------------------------------------
B200LP_API int b200lp_shard_reset(b200lp_solver* s, int64_t max_pivots) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(ensure_hist(s, std::min<int64_t>(max_pivots, (int64_t)1 << 20)));
    CKR(launch_reset(s, max_pivots, false));
    return 0;
}

B200LP_API int b200lp_shard_candidate(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, double* cand_dev) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!cand_dev) return fail(B200LP_E_INVALID, "cand_dev is NULL");
    CKR(check_opts(o));
    CKR(set_device(s));
    CKR(launch_price(s, obj_row, o->rule, o->eps_cost, true));
    const int blocks = clampi((s->R + PICK_THREADS - 1) / PICK_THREADS, 1, 2 * s->sm_count);
    k_shard_extract<<<blocks, PICK_THREADS, 0, s->stream>>>(s->T, s->R, s->ld, s->st.p, cand_dev);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

B200LP_API int b200lp_shard_pivot(b200lp_solver* s, const b200lp_opts* o, const double* gathered_dev, int32_t world,
                                  int32_t rank) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!gathered_dev || world < 1 || rank < 0 || rank >= world) return fail(B200LP_E_INVALID, "bad shard arguments");
    CKR(check_opts(o));
    CKR(set_device(s));
    const int64_t stride = s->R + 2;
    k_shard_winner<<<1, 32, 0, s->stream>>>(gathered_dev, stride, world, rank, o->rule == B200LP_RULE_BLAND, s->st.p);
    s->launches++;
    CK(cudaGetLastError());
    CKR(launch_ratio(s, o->eps_pivot, false, gathered_dev, stride));
    CKR(launch_update(s, o->update_variant));
    return 0;
}

// ---- look-ahead (blocked) loop on a column shard: same exchange per pivot, one tableau pass per K pivots ----
B200LP_API int b200lp_shard_blk_begin(b200lp_solver* s, int64_t obj_row) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (obj_row < s->m || obj_row >= s->R) return fail(B200LP_E_INVALID, "obj_row %lld is not an objective row", (long long)obj_row);
    CKR(set_device(s));
    CKR(launch_flush(s));
    CKR(blk_setup(s));
    const int ib = clampi((std::max(s->R, s->C) + BLK_THREADS - 1) / BLK_THREADS, 1, 2 * s->sm_count);
    k_blk_init<<<ib, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->C, s->ld, obj_row, s->st.p, s->blk);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

B200LP_API int b200lp_shard_blk_candidate(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, double* cand_dev) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!cand_dev) return fail(B200LP_E_INVALID, "cand_dev is NULL");
    CKR(check_opts(o));
    CKR(set_device(s));
    CKR(launch_blk_rowprice(s, o, obj_row, true));
    const int blocks = clampi((s->R + BLK_THREADS - 1) / BLK_THREADS, 1, 2 * s->sm_count);
    k_blk_shard_extract<<<blocks, BLK_THREADS, 0, s->stream>>>(s->T, s->R, s->ld, s->st.p, s->blk, cand_dev);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

B200LP_API int b200lp_shard_blk_pivot(b200lp_solver* s, const b200lp_opts* o, const double* gathered_dev, int32_t world,
                                      int32_t rank) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!gathered_dev || world < 1 || rank < 0 || rank >= world) return fail(B200LP_E_INVALID, "bad shard arguments");
    CKR(check_opts(o));
    CKR(set_device(s));
    const int64_t stride = s->R + 2;
    k_shard_winner<<<1, 32, 0, s->stream>>>(gathered_dev, stride, world, rank, o->rule == B200LP_RULE_BLAND, s->st.p);
    s->launches++;
    CK(cudaGetLastError());
    CKR(launch_blk_ratio(s, o, gathered_dev, stride));
    return 0;
}

B200LP_API int b200lp_shard_blk_flush(b200lp_solver* s, int64_t obj_row) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    CKR(enqueue_blk_flush(s, BLK_KMAX, obj_row));
    return 0;
}

// ---- peer-memory exchange: one fused kernel per pivot decision (kernels_shard.cuh) ----
static int64_t p2p_xstride(int64_t R) { return (R + 3) & ~(int64_t)1; }  // even: 16-byte stores into every record

B200LP_API int64_t b200lp_p2p_bytes(int64_t R, int32_t world) {
    if (R < 1 || world < 1 || world > 16) return -1;
    return (2 * (int64_t)world * p2p_xstride(R) + 2 * (int64_t)world) * 8;
}

B200LP_API int b200lp_p2p_connect(b200lp_solver* s, void* const* bases, int32_t world, int32_t rank) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!bases || world < 1 || world > 16 || rank < 0 || rank >= world) return fail(B200LP_E_INVALID, "bad peer arguments");
    if (!s->cluster_ctas) return fail(B200LP_E_STATE, "thread-block clusters are not available: use the all-gather exchange");
    CKR(set_device(s));
    memset(&s->p2p, 0, sizeof(s->p2p));
    for (int g = 0; g < world; ++g) {
        if (!bases[g]) return fail(B200LP_E_INVALID, "peer %d has no region", g);
        if (((uintptr_t)bases[g]) & 15) return fail(B200LP_E_INVALID, "region of peer %d is not 16-byte aligned", g);
        s->p2p.base[g] = (double*)bases[g];
    }
    s->p2p.world = world;
    s->p2p.rank = rank;
    s->p2p.xstride = p2p_xstride(s->R);
    // own flags and the generation counter start at 0 (the caller synchronises all ranks after connect, before the
    // first exchange); from here on the counter is never reset, whatever happens to the tableau or the loop state
    CKR(s->xgen.ensure(1));
    CKR(s->shard_ctx.ensure(16));
    CK(cudaMemsetAsync(s->p2p.base[rank] + 2 * (int64_t)world * s->p2p.xstride, 0, (size_t)2 * world * 8, s->stream));
    CK(cudaMemsetAsync(s->xgen.p, 0, sizeof(long long), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->shard_ctx_host.clear();
    s->p2p_on = true;
    s->bind_epoch++;
    return 0;
}

// A number that changes whenever anything a caller-captured launch of this solver bakes in may have changed: workspace
// buffers (re)allocated -- history arrays growing with max_pivots, look-ahead buffers, the pivot-column copy --, the
// tableau re-attached, snapshots or peer regions re-bound.  Callers that replay their own CUDA graphs of b200lp_shard_*
// launches key them by it (the library's own graphs are keyed by the baked values themselves, GraphKey).
B200LP_API int b200lp_binding_epoch(b200lp_solver* s, int64_t* epoch) {
    if (!s || !epoch) return fail(B200LP_E_INVALID, "solver or epoch is NULL");
    *epoch = (int64_t)(g_alloc_epoch.load(std::memory_order_relaxed) + s->bind_epoch);
    return 0;
}

static ShardCtx make_shard_ctx(b200lp_solver* s) {
    ShardCtx X;
    memset(&X, 0, sizeof(X));
    X.A.T = s->T;
    X.A.R = s->R;
    X.A.m = s->m;
    X.A.C = s->C;
    X.A.ld = s->ld;
    X.A.rowlab = s->rowlab.p;
    X.A.collab = s->collab.p;
    X.A.art_base = s->art_base;
    X.A.st = s->st.p;
    X.A.col = s->col.p;
    X.A.B = s->blk;
    X.A.h_row = s->h_row.p;
    X.A.h_col = s->h_col.p;
    X.A.h_enter = s->h_enter.p;
    X.A.h_leave = s->h_leave.p;
    X.A.hist_cap = s->hist_cap;
    X.P = s->p2p;
    X.xgen = s->xgen.p;
    X.error = s->pk_err.p;
    return X;
}

// Upload the contexts of `n` shards into ss[0]'s context array when they differ from what is there (never inside a graph
// capture in practice: the first chunk of a run is enqueued eagerly, its capture repeats the same arguments).
static int sync_shard_ctx(b200lp_solver* const* ss, int n) {
    std::vector<ShardCtx> want((size_t)n);
    for (int k = 0; k < n; ++k) want[(size_t)k] = make_shard_ctx(ss[k]);
    b200lp_solver* s0 = ss[0];
    if (s0->shard_ctx_host.size() == (size_t)n && memcmp(s0->shard_ctx_host.data(), want.data(), sizeof(ShardCtx) * n) == 0) return 0;
    CKR(s0->shard_ctx.ensure((size_t)std::max(n, 16)));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (s0->stream != (cudaStream_t)0) CK(cudaStreamIsCapturing(s0->stream, &cap));
    if (cap != cudaStreamCaptureStatusNone)
        return fail(B200LP_E_STATE, "the shard context changed inside a stream capture: run one pivot eagerly first");
    CK(cudaStreamSynchronize(s0->stream));
    CK(cudaMemcpy(s0->shard_ctx.p, want.data(), sizeof(ShardCtx) * n, cudaMemcpyHostToDevice));
    s0->shard_ctx_host = want;
    return 0;
}

static int launch_shard_pick(b200lp_solver* s0, int n_ctx, const b200lp_opts* o, int64_t obj_row, bool lookahead,
                             bool pre = false) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(s0->cluster_ctas * n_ctx));
    cfg.blockDim = dim3(CL_THREADS);
    cfg.stream = s0->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = s0->cluster_ctas;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;  // emulation: clusters that wait for one another must all be resident
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    const ShardCtx* ctxs = s0->shard_ctx.p;
    const bool bland = o->rule == B200LP_RULE_BLAND;
    auto launch = [&]() -> cudaError_t {
        if (lookahead && pre)
            return bland ? cudaLaunchKernelEx(&cfg, k_shard_pick<true, true, true>, ctxs, obj_row, o->eps_cost, o->eps_pivot)
                         : cudaLaunchKernelEx(&cfg, k_shard_pick<false, true, true>, ctxs, obj_row, o->eps_cost, o->eps_pivot);
        if (lookahead) return bland ? cudaLaunchKernelEx(&cfg, k_shard_pick<true, true>, ctxs, obj_row, o->eps_cost, o->eps_pivot)
                                    : cudaLaunchKernelEx(&cfg, k_shard_pick<false, true>, ctxs, obj_row, o->eps_cost, o->eps_pivot);
        return bland ? cudaLaunchKernelEx(&cfg, k_shard_pick<true, false>, ctxs, obj_row, o->eps_cost, o->eps_pivot)
                     : cudaLaunchKernelEx(&cfg, k_shard_pick<false, false>, ctxs, obj_row, o->eps_cost, o->eps_pivot);
    };
    cfg.numAttrs = n_ctx > 1 ? 2 : 1;
    cudaError_t e = launch();
    if (e != cudaSuccess && n_ctx > 1) {  // a driver that refuses cooperative cluster launches: the clusters of a test
        cudaGetLastError();               // (a few dozen CTAs on an idle GPU) are co-resident anyway
        cfg.numAttrs = 1;
        e = launch();
    }
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, "fused shard pick launch failed: %s", cudaGetErrorString(e));
    s0->launches++;
    return 0;
}

static int check_shard_args(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    if (!s->p2p_on) return fail(B200LP_E_STATE, "b200lp_p2p_connect was not called");
    if (s->p2p.xstride != p2p_xstride(s->R)) return fail(B200LP_E_STATE, "the tableau changed its row count after b200lp_p2p_connect");
    if (obj_row < s->m || obj_row >= s->R) return fail(B200LP_E_INVALID, "obj_row %lld is not an objective row", (long long)obj_row);
    return check_opts(o);
}

// Look-ahead decision, first two launches of three: row part + pricing on all SMs (k_blk_rowprice), then the candidate
// column gathered, replayed and pushed to every peer on all SMs (k_blk_shard_push); k_shard_pick<.., true, true> follows.
// `ctxs` / `index`: the context array the push kernel reads and this shard's entry in it.
static int launch_shard_push_chain(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, const ShardCtx* ctxs, int index) {
    CKR(launch_blk_rowprice(s, o, obj_row, true));
    const int blocks = clampi((s->R + 2 * BLK_THREADS - 1) / (2 * BLK_THREADS), 1, 2 * s->sm_count);
    k_blk_shard_push<<<blocks, BLK_THREADS, 0, s->stream>>>(ctxs, index);
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

// One pivot of the sharded loop with the peer-memory exchange.  Rank-1 loop: fused pick + update.  Look-ahead loop: the
// decision alone, as rowprice -> push -> pick (three launches, the heavy parts on all SMs) unless the solver was
// created with B200LP_SHARD_LA_FUSED set (diagnostic: everything inside the one-cluster kernel).
B200LP_API int b200lp_shard_fused(b200lp_solver* s, const b200lp_opts* o, int64_t obj_row, int32_t lookahead) {
    CKR(check_shard_args(s, o, obj_row));
    CKR(set_device(s));
    b200lp_solver* one[1] = {s};
    CKR(sync_shard_ctx(one, 1));
    const bool pre = lookahead && !s->shard_la_fused;
    if (pre) CKR(launch_shard_push_chain(s, o, obj_row, s->shard_ctx.p, 0));
    CKR(launch_shard_pick(s, 1, o, obj_row, lookahead != 0, pre));
    if (!lookahead) CKR(launch_update(s, o->update_variant));
    return 0;
}

// Single-GPU emulation of `n` shards (tests): ONE launch runs one cluster per shard, all resident at once, so the shards'
// waits for one another are satisfied inside the kernel.  All solvers must live on the same device and stream.
B200LP_API int b200lp_shard_fused_multi(b200lp_solver* const* ss, int32_t n, const b200lp_opts* o, int64_t obj_row,
                                        int32_t lookahead) {
    if (!ss || n < 1 || n > 8) return fail(B200LP_E_INVALID, "bad shard list");
    for (int k = 0; k < n; ++k) {
        CKR(check_shard_args(ss[k], o, obj_row));
        if (ss[k]->device != ss[0]->device || ss[k]->stream != ss[0]->stream)
            return fail(B200LP_E_INVALID, "emulated shards must share one device and one stream");
        if (ss[k]->p2p.world != n || ss[k]->p2p.rank != k) return fail(B200LP_E_INVALID, "shard %d is not rank %d of %d", k, k, (int)n);
    }
    CKR(set_device(ss[0]));
    CKR(sync_shard_ctx(ss, n));
    const bool pre = lookahead && !ss[0]->shard_la_fused;
    if (pre)  // pushes wait for nobody: the shards' chains run one after the other, then ONE launch with all the clusters
        for (int k = 0; k < n; ++k) CKR(launch_shard_push_chain(ss[k], o, obj_row, ss[0]->shard_ctx.p, k));
    CKR(launch_shard_pick(ss[0], n, o, obj_row, lookahead != 0, pre));
    if (!lookahead)
        for (int k = 0; k < n; ++k) CKR(launch_update(ss[k], o->update_variant));
    return 0;
}

B200LP_API int b200lp_shard_state(b200lp_solver* s, int32_t* done, int32_t* status, int64_t* n_pivots) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    DevState st;
    CKR(read_state(s, &st));
    if (st.done && st.status == B200LP_STATUS_NUMERICAL && s->p2p_on) {
        int32_t err = 0;
        CK(cudaMemcpy(&err, s->pk_err.p, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CK(cudaMemset(s->pk_err.p, 0, sizeof(err)));
            return fail(B200LP_E_CUDA, "peer-memory exchange: a peer's candidate never arrived (rank missing or out of step)");
        }
    }
    if (done) *done = st.done;
    if (status) *status = st.status;
    if (n_pivots) *n_pivots = st.n_pivots;
    return 0;
}

B200LP_API int b200lp_read_history(b200lp_solver* s, int64_t cap, int32_t* piv_row, int32_t* piv_col, int32_t* enter_lab,
                                   int32_t* leave_lab, int64_t* n_out) {
    if (!s || !s->T) return fail(B200LP_E_STATE, "no tableau bound");
    CKR(set_device(s));
    DevState st;
    CKR(read_state(s, &st));
    b200lp_result r;
    memset(&r, 0, sizeof(r));
    r.piv_row = piv_row;
    r.piv_col = piv_col;
    r.enter_lab = enter_lab;
    r.leave_lab = leave_lab;
    r.hist_cap = cap;
    CKR(copy_history(s, &r, st.n_pivots));
    if (n_out) *n_out = std::min<int64_t>(std::min<int64_t>(st.n_pivots, cap), s->hist_cap);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// batched small LPs
// ------------------------------------------------------------------------------------------------------
B200LP_API int b200lp_solve_batched(b200lp_solver* s, int64_t B, int64_t m, int64_t n, const double* A, const double* b,
                                    const double* c, const int8_t* ops, const b200lp_opts* o, int32_t* status, double* fun,
                                    double* x, int32_t* n_pivots, int32_t* piv_log, int64_t log_cap, int32_t on_device,
                                    double* device_ms) {
    if (!s) return fail(B200LP_E_INVALID, "solver is NULL");
    CKR(check_opts(o));
    if (B < 0 || m < 0 || n < 1) return fail(B200LP_E_INVALID, "bad batch shape B=%lld m=%lld n=%lld", (long long)B, (long long)m, (long long)n);
    if (B == 0) return 0;
    if (B > 0x7fffffff) return fail(B200LP_E_INVALID, "batch of %lld LPs: split it (the work counter of a launch is 32 bits wide)", (long long)B);
    if (!A && m > 0) return fail(B200LP_E_INVALID, "A is NULL");
    if (!c || (m > 0 && (!b || !ops)) || !status || !fun || !n_pivots) return fail(B200LP_E_INVALID, "NULL argument");
    CKR(set_device(s));

    BatchedParams P;
    memset(&P, 0, sizeof(P));
    int64_t ld = n + m + 1;
    if ((ld & 1) == 0) ++ld;
    const int64_t R = m + 2;
    size_t warp_bytes = (size_t)(R * ld + R) * 8 + (size_t)(R + ld) * 4;
    warp_bytes = (warp_bytes + 15) / 16 * 16;
    const size_t smem_max = BATCHED_SMEM_MAX;
    if (warp_bytes > smem_max)
        return fail(B200LP_E_INVALID, "LP of %lld x %lld needs %zu bytes of shared memory per warp; use b200lp_solve_dense",
                    (long long)m, (long long)n, warp_bytes);
    // warps per CTA: the warps are persistent (they draw LPs from a counter), so the CTA size only decides how many warps
    // fit one SM's shared memory -- take the size with the most resident warps (ties: the larger CTA).  The plan depends
    // on the LP shape only and is kept: the eight occupancy queries cost as much as a small batch.
    if (s->bplan.m != m || s->bplan.n != n) {
        int wpc_best = 1, per_sm_best = 0, best_warps = 0;
        for (int cand = 1; cand <= 8 && (size_t)cand * warp_bytes <= smem_max; ++cand) {
            int occ = 0;
            if (m + 2 >= 16) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_solve_batched<4>, cand * 32, (size_t)cand * warp_bytes));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_solve_batched<1>, cand * 32, (size_t)cand * warp_bytes));
            if (occ * cand >= best_warps) {
                best_warps = occ * cand;
                wpc_best = cand;
                per_sm_best = occ;
            }
        }
        s->bplan.m = m;
        s->bplan.n = n;
        s->bplan.wpc = wpc_best;
        s->bplan.per_sm = per_sm_best;
    }
    const int wpc = s->bplan.wpc, per_sm = s->bplan.per_sm;
    const size_t smem = (size_t)wpc * warp_bytes;

    const double *dA = A, *db = b, *dc = c;
    const int8_t* dops = ops;
    int32_t *dstatus = status, *dnp = n_pivots, *dlog = piv_log;
    double *dfun = fun, *dx = x;
    const bool want_log = piv_log && log_cap > 0;
    if (!on_device) {
        CKR(s->sA.ensure((size_t)std::max<int64_t>(B * m * n, 1)));
        CKR(s->sb.ensure((size_t)std::max<int64_t>(B * m, 1)));
        CKR(s->sc.ensure((size_t)(B * n)));
        CKR(s->sops.ensure((size_t)std::max<int64_t>(B * m, 1)));
        CKR(s->sstatus.ensure((size_t)B));
        CKR(s->snpiv.ensure((size_t)B));
        CKR(s->sfun.ensure((size_t)B));
        if (x) CKR(s->sx.ensure((size_t)(B * n)));
        if (want_log) CKR(s->slog.ensure((size_t)(B * log_cap * 2)));
        dA = s->sA.p;
        db = s->sb.p;
        dc = s->sc.p;
        dops = s->sops.p;
        dstatus = s->sstatus.p;
        dnp = s->snpiv.p;
        dfun = s->sfun.p;
        dx = x ? s->sx.p : nullptr;
        dlog = want_log ? s->slog.p : nullptr;
    }
    P.m = (int32_t)m;
    P.n = (int32_t)n;
    P.ld = (int32_t)ld;
    P.rule = o->rule;
    P.auto_budget = o->max_pivots >= AUTO_BUDGET ? 1 : 0;
    P.max_pivots = (int32_t)std::min<int64_t>(o->max_pivots, 0x3fffffff);  // automatic budgets are sized per LP in the kernel
    P.log_cap = (int32_t)(dlog ? log_cap : 0);
    P.eps_cost = o->eps_cost;
    P.eps_pivot = o->eps_pivot;
    P.eps_feas = o->eps_feas;
    P.warp_bytes = warp_bytes;

    // Host inputs: the batch is cut into chunks that alternate between two streams, so that the H2D copy of one
    // chunk overlaps the kernel of the previous one and the D2H of the one before (PCIe is the bound of this path).
    // Chunks are sized by BYTES (about 6 MB of input each, 2 .. 16 chunks): a rank's share of a sharded batch is small in
    // LPs but still tens of MB, and without chunks its kernel waits for the whole upload.
    int nchunk = 1;
    if (!on_device) {
        const double in_bytes = (double)B * ((double)m * n * 8 + (double)m * 9 + (double)n * 8);
        nchunk = (int)std::max(1.0, std::min((double)BATCHED_MAX_CHUNKS, in_bytes / 6.0e6));
        if (nchunk > B) nchunk = (int)B;
    }
    cudaStream_t q[2] = {s->stream, nchunk > 1 ? s->stream2 : s->stream};
    // persistent warps: one grid of resident CTAs per launch, LPs drawn from a counter
    const int64_t resident_ctas = (int64_t)std::max(per_sm, 1) * s->sm_count;
    CK(cudaMemsetAsync(s->snext.p, 0, BATCHED_MAX_CHUNKS * sizeof(unsigned int), s->stream));
    CK(cudaEventRecord(s->ev0, s->stream));
    if (nchunk > 1) CK(cudaStreamWaitEvent(s->stream2, s->ev0, 0));
    for (int k = 0; k < nchunk; ++k) {
        const int64_t lo = B * k / nchunk, hi = B * (k + 1) / nchunk, nb = hi - lo;
        if (nb <= 0) continue;
        cudaStream_t st = q[k & 1];
        if (!on_device) {
            if (m > 0) {
                CK(cudaMemcpyAsync(s->sA.p + lo * m * n, A + lo * m * n, (size_t)(nb * m * n) * 8, cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(s->sb.p + lo * m, b + lo * m, (size_t)(nb * m) * 8, cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(s->sops.p + lo * m, ops + lo * m, (size_t)(nb * m), cudaMemcpyHostToDevice, st));
            }
            CK(cudaMemcpyAsync(s->sc.p + lo * n, c + lo * n, (size_t)(nb * n) * 8, cudaMemcpyHostToDevice, st));
        }
        P.B = nb;
        P.A = dA + lo * m * n;
        P.b = db + lo * m;
        P.c = dc + lo * n;
        P.ops = dops + lo * m;
        P.status = dstatus + lo;
        P.fun = dfun + lo;
        P.x = dx ? dx + lo * n : nullptr;
        P.n_pivots = dnp + lo;
        P.piv_log = dlog ? dlog + lo * log_cap * 2 : nullptr;
        P.next = s->snext.p + k;
        const int64_t blocks = std::min<int64_t>((nb + wpc - 1) / wpc, resident_ctas);
        if (m + 2 >= 16) k_solve_batched<4><<<(unsigned)blocks, wpc * 32, smem, st>>>(P);
        else k_solve_batched<1><<<(unsigned)blocks, wpc * 32, smem, st>>>(P);
        s->launches++;
        CK(cudaGetLastError());
        if (!on_device) {
            CK(cudaMemcpyAsync(status + lo, dstatus + lo, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(n_pivots + lo, dnp + lo, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(fun + lo, dfun + lo, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
            if (x) CK(cudaMemcpyAsync(x + lo * n, dx + lo * n, (size_t)(nb * n) * 8, cudaMemcpyDeviceToHost, st));
            if (dlog) CK(cudaMemcpyAsync(piv_log + lo * log_cap * 2, dlog + lo * log_cap * 2, (size_t)(nb * log_cap * 2) * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    if (nchunk > 1) {
        CK(cudaEventRecord(s->ev_state[0], s->stream2));
        CK(cudaStreamWaitEvent(s->stream, s->ev_state[0], 0));
    }
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (device_ms) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        *device_ms = ms;
    }
    return 0;
}
