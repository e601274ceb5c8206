// kernels_onchip.cuh -- the pivot loop for tableaux that fit the chip's aggregate shared memory (~30 MB):
// one persistent cooperative kernel, the tableau column-sharded over the SMs and RESIDENT IN SHARED MEMORY for the
// whole phase, ONE grid barrier per pivot.  This is the path of BASELINE config 2 (1024 x 1024, 8.4 MB), which is
// latency-bound, not HBM-bound: the multi-kernel loop costs 12-16 us per pivot there, this loop 4.7 us.
//
// Layout (the multi-GPU column sharding of sharded.py, applied across SMs): CTA g owns columns [lo_g, hi_g) of EVERY
// row plus its own replica of the right-hand-side column and of the row labels.  Per pivot:
//   1. local argmin over the CTA's slice of the objective row;
//   2. publish [reduced cost, variable id, candidate column (R doubles)] to a global exchange buffer (L2 resident);
//   3. grid barrier (the only one);
//   4. every CTA reads the G headers, takes the same decision (total order), copies the winner's column, runs the
//      ratio test redundantly on its RHS replica, scales its slice of the pivot row and applies the rank-1 update
//      to its columns in shared memory.
// No pivot-row exchange is needed under column sharding, and the ratio test needs no second reduction because the RHS
// is replicated.  Arithmetic and tie-breaking are those of DESIGN.md section 2: bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace b200lp {

#ifndef B200LP_ONCHIP_THREADS
#define B200LP_ONCHIP_THREADS 512
#endif
constexpr int ONCHIP_THREADS = B200LP_ONCHIP_THREADS;
constexpr int ONCHIP_WARPS = ONCHIP_THREADS / 32;
static_assert(3 * ONCHIP_WARPS * 16 <= 1024, "static shared memory of k_solve_onchip exceeds the 1 KB set aside for it");
constexpr int ONCHIP_CHUNK = 8;  // columns of a row held in registers at once by the update
// stride of one candidate record [cost, id, column (R doubles)] in the exchange buffer: even, so that the 16-byte
// header and the column behind it can be read with 128-bit loads
__host__ __device__ inline int64_t onchip_xstride(int64_t R) { return (R + 3) & ~(int64_t)1; }
// shared-memory budget: the 227 KB a CTA can opt in to, minus the kernel's static shared memory
constexpr size_t ONCHIP_SMEM_MAX = 232448 - 1024;

struct OnchipParams {
    double* T;           // global tableau (loaded at start, written back at the end)
    int64_t R, m, C, ld;
    int64_t obj_row;
    int32_t* rowlab;     // global labels (read at start, written back at the end)
    int32_t* collab;
    int32_t art_base;
    int32_t rule;        // 0 Dantzig, 1 Bland
    double eps_cost, eps_pivot;
    DevState* st;
    double* xbuf;        // exchange: 2 x G x (R + 2) doubles
    unsigned long long* barrier;  // monotonic arrival counter (zeroed before the launch)
    int32_t* h_row;
    int32_t* h_col;
    int32_t* h_enter;
    int32_t* h_leave;
    int64_t hist_cap;
    int32_t stride;      // odd row stride of the shared-memory slice, >= widest slice + 1
    int32_t* error;      // set to 1 when a grid barrier times out (a CTA went missing): the host reports it
    long long time_budget_ns;  // wall-clock bound of this launch on %globaltimer; 0 = none
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// header id of CTA 0 when the wall-clock bound has expired: every CTA reads the G headers, so all of them stop at the
// same pivot (a time check per CTA could split the grid)
#define ONCHIP_TIMEOUT_ID (-2.0)

// Grid barrier on a monotonic arrival counter (the kernel is launched cooperatively, so all CTAs are resident).
// One thread per CTA arrives with a release reduction (no return value to wait for) and polls with acquire loads.
// (A flag-per-CTA all-to-all exchange without a counter was measured slower on B200: 148 x 148 acquire-polling
// threads cost more than one poller per CTA plus a separate read of the headers.)
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ONCHIP_THREADS, 1) k_solve_onchip(const OnchipParams P) {
    extern __shared__ __align__(16) uint8_t smem_onchip[];
    __shared__ Key sk_price[ONCHIP_WARPS], sk_decide[ONCHIP_WARPS], sk_ratio[ONCHIP_WARPS];
    __shared__ int wd_failed;  // this CTA's poll of the grid barrier gave up
    const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x;
    const int64_t R = P.R, m = P.m, C = P.C;
    const int stride = P.stride;
    const int64_t lo = (C - 1) * g / G, hi = (C - 1) * (g + 1) / G;
    const int wl = (int)(hi - lo);  // local real columns; local column wl = RHS replica

    double* Tl = reinterpret_cast<double*>(smem_onchip);
    double* colbuf = Tl + (size_t)R * stride;
    double* qloc = colbuf + R;
    int32_t* rl = reinterpret_cast<int32_t*>(qloc + stride);
    int32_t* cl = rl + R;

    // ---- load the slice ----
    for (int64_t e = tid; e < R * (wl + 1); e += ONCHIP_THREADS) {
        const int64_t i = e / (wl + 1);
        const int j = (int)(e - i * (wl + 1));
        Tl[i * stride + j] = P.T[i * P.ld + (j < wl ? lo + j : C - 1)];
    }
    for (int64_t i = tid; i < R; i += ONCHIP_THREADS) rl[i] = P.rowlab[i];
    for (int j = tid; j < wl; j += ONCHIP_THREADS) cl[j] = P.collab[lo + j];
    __syncthreads();

    long long n_pivots = P.st->n_pivots;
    const long long max_pivots = P.st->max_pivots;
    int status = -1;
    const int64_t xstride = onchip_xstride(R);
    const int64_t obj = P.obj_row;
    unsigned long long bar_round = 0;

    // The update of a pivot is applied in three parts so that most of it runs while the grid barrier of the NEXT pivot is
    // in flight: (a) the objective row, which is all the next pricing reads; (b) the column that pricing then picks, which
    // is what gets published; -- arrive at the barrier -- (c) everything else, before the wait.  Every element still
    // receives exactly one fma with the operands of DESIGN.md section 2, so the bits do not change.
    const unsigned long long t_start = (g == 0 && tid == 0 && P.time_budget_ns > 0) ? global_timer_ns() : 0ull;
    if (tid == 0) wd_failed = 0;
    bool pending = false;  // colbuf / qloc / pr / ps describe a pivot whose update has not been applied yet
    int pr = -1, ps = -1;  // its row and, in the CTA that owns the entering column, its local column (else -1)

    // rows of the pending pivot other than row pr and the objective row, columns [0, wl] except `skip`
    auto update_rest = [&](const int skip) {
        for (int j0 = 0; j0 <= wl; j0 += ONCHIP_CHUNK) {
            double q[ONCHIP_CHUNK];
            bool act[ONCHIP_CHUNK];
#pragma unroll
            for (int u = 0; u < ONCHIP_CHUNK; ++u) {
                act[u] = (j0 + u <= wl) && (j0 + u != skip);
                q[u] = act[u] ? qloc[j0 + u] : 0.0;
            }
            for (int64_t i = tid; i < R; i += ONCHIP_THREADS) {
                if (i == pr || i == obj) continue;
                const double nc = -colbuf[i];
                double* row = Tl + i * stride + j0;
                double t[ONCHIP_CHUNK];
#pragma unroll
                for (int u = 0; u < ONCHIP_CHUNK; ++u) t[u] = (act[u] && j0 + u != ps) ? row[u] : 0.0;
#pragma unroll
                for (int u = 0; u < ONCHIP_CHUNK; ++u)
                    if (act[u]) row[u] = __fma_rn(nc, q[u], t[u]);
            }
        }
        for (int j = tid; j <= wl; j += ONCHIP_THREADS) Tl[(int64_t)pr * stride + j] = qloc[j];  // row pr was skipped above
    };
    // the objective row of the pending pivot
    auto update_obj_row = [&]() {
        const double nc = -colbuf[obj];
        for (int j = tid; j <= wl; j += ONCHIP_THREADS) {
            const double t = (j == ps) ? 0.0 : Tl[obj * stride + j];
            Tl[obj * stride + j] = __fma_rn(nc, qloc[j], t);
        }
    };

    for (long long it = 0;; ++it) {
        if (n_pivots >= max_pivots) {
            if (pending) {
                update_obj_row();
                update_rest(-1);
            }
            status = 1;
            break;
        }
        // ---- 1. local pricing on the current objective row ----
        if (pending) {
            update_obj_row();
            __syncthreads();
        }
        Key k = key_none();
        for (int j = wl <= 32 ? (tid & 31) : tid; j < wl; j += ONCHIP_THREADS) {
            const int32_t lab = cl[j];
            const double v = Tl[obj * stride + j];
            if (lab < P.art_base && v < -P.eps_cost) {
                Key c;
                c.v = v;
                c.lab = lab;
                c.pos = j;
                k = P.rule ? key_min<true>(k, c) : key_min<false>(k, c);
            }
        }
        // (a slice of at most 32 columns is priced by every warp on its own: no shared-memory round trip, no barrier)
        const Key mine = wl <= 32 ? (P.rule ? warp_key_min<true>(k) : warp_key_min<false>(k))
                                  : (P.rule ? block_key_min_all<true>(k, sk_price) : block_key_min_all<false>(k, sk_price));
        // ---- 2. publish the candidate (bringing its column up to date on the way) ----
        double* slot = P.xbuf + ((size_t)(it & 1) * G + g) * xstride;
        if (tid == 0) {
            slot[0] = mine.v;
            slot[1] = mine.lab == B200LP_NO_LAB ? -1.0 : (double)mine.lab;
            if (g == 0 && P.time_budget_ns > 0 && (it & 15) == 0 &&
                global_timer_ns() - t_start > (unsigned long long)P.time_budget_ns)
                slot[1] = ONCHIP_TIMEOUT_ID;
        }
        const int cand = mine.lab != B200LP_NO_LAB ? mine.pos : -1;
        if (cand >= 0) {
            if (pending) {
                const double qc = qloc[cand];
                for (int64_t i = tid; i < R; i += ONCHIP_THREADS) {
                    double* e = Tl + i * stride + cand;
                    double v;
                    if (i == pr) v = qc;
                    else if (i == obj) v = *e;
                    else v = __fma_rn(-colbuf[i], qc, cand == ps ? 0.0 : *e);
                    *e = v;
                    slot[2 + i] = v;
                }
            } else {
                for (int64_t i = tid; i < R; i += ONCHIP_THREADS) slot[2 + i] = Tl[i * stride + cand];
            }
        }
        // ---- 3. the grid barrier of this pivot: arrive, finish the pending update, wait ----
        ++bar_round;
        __syncthreads();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(P.barrier) : "memory");
        if (pending) {
            update_rest(cand);
            pending = false;
        }
        if (tid == 0) {
            const unsigned long long target = bar_round * (unsigned long long)G;
            unsigned long long v;
            unsigned int spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(P.barrier) : "memory");
                // watchdog: a barrier that cannot complete must not hang the GPU.  Polls are counted (an integer add per
                // round trip to L2, ~0.4 us) rather than timed: reading the clock between polls delays the exit
                if (v < target && ++spins > (1u << 23)) {  // a few seconds
                    wd_failed = 1;
                    *P.error = 1;
                    break;
                }
            } while (v < target);
        }
        __syncthreads();
        if (wd_failed) {
            status = 4;
            break;
        }
        // ---- 4. global decision (identical on every CTA): one 16-byte header per candidate ----
        k = key_none();
        const double* xb = P.xbuf + (size_t)(it & 1) * G * xstride;
        for (int b = tid; b < G; b += ONCHIP_THREADS) {
            const double2 h = __ldcg(reinterpret_cast<const double2*>(xb + (size_t)b * xstride));
            if (h.y == ONCHIP_TIMEOUT_ID) {  // wins under both orders: lowest id, lowest value
                k.v = __longlong_as_double((long long)0xfff0000000000000ull);  // -inf
                k.lab = -1;
                k.pos = b;
            } else if (h.y >= 0.0) {
                Key c;
                c.v = h.x;
                c.lab = (int32_t)h.y;
                c.pos = b;
                k = P.rule ? key_min<true>(k, c) : key_min<false>(k, c);
            }
        }
        const Key win = P.rule ? block_key_min_all<true>(k, sk_decide) : block_key_min_all<false>(k, sk_decide);
        if (win.lab == B200LP_NO_LAB) {
            status = 0;
            break;
        }
        if (win.lab < 0) {  // wall-clock bound expired (nothing is pending here: the tableau is consistent)
            status = 1;
            break;
        }
        const double2* wcol = reinterpret_cast<const double2*>(xb + (size_t)win.pos * xstride + 2);
        for (int64_t i2 = tid; 2 * i2 < R; i2 += ONCHIP_THREADS) {
            const double2 v = __ldcg(wcol + i2);  // the pad element behind an odd R belongs to the record
            colbuf[2 * i2] = v.x;
            if (2 * i2 + 1 < R) colbuf[2 * i2 + 1] = v.y;
        }
        __syncthreads();
        // ---- ratio test on the RHS replica: two rows per trip, so that their divisions overlap ----
        k = key_none();
        for (int64_t i0 = tid; i0 < m; i0 += 2 * ONCHIP_THREADS) {
            const int64_t i1 = i0 + ONCHIP_THREADS;
            const bool in1 = i1 < m;
            const int32_t lab0 = rl[i0], lab1 = in1 ? rl[i1] : -1;
            const double a0 = colbuf[i0], a1 = in1 ? colbuf[i1] : 0.0;
            const bool ok0 = lab0 >= 0 && a0 > P.eps_pivot, ok1 = lab1 >= 0 && a1 > P.eps_pivot;
            const double b0 = Tl[i0 * stride + wl], b1 = in1 ? Tl[i1 * stride + wl] : 0.0;
            const double r0 = b0 / (ok0 ? a0 : 1.0), r1 = b1 / (ok1 ? a1 : 1.0);
            if (ok0) {
                Key c;
                c.v = r0;
                c.lab = lab0;
                c.pos = (int32_t)i0;
                k = key_min<false>(k, c);
            }
            if (ok1) {
                Key c;
                c.v = r1;
                c.lab = lab1;
                c.pos = (int32_t)i1;
                k = key_min<false>(k, c);
            }
        }
        const int r = block_key_min_all<false>(k, sk_ratio).pos;
        if (r < 0) {
            status = 3;
            break;
        }
        // ---- the pivot is decided: scaled pivot row, labels, history; the update itself stays pending ----
        const int s_local = (win.pos == g) ? cand : -1;
        const double p = colbuf[r];
        const double inv_p = 1.0 / p;
        for (int j = tid; j <= wl; j += ONCHIP_THREADS) qloc[j] = (j == s_local) ? inv_p : Tl[(int64_t)r * stride + j] / p;
        if (tid == 0) {
            const int32_t leave = rl[r];
            rl[r] = win.lab;
            if (s_local >= 0) cl[s_local] = leave;
            if (n_pivots < P.hist_cap) {
                if (g == 0) {
                    P.h_row[n_pivots] = r;
                    P.h_enter[n_pivots] = win.lab;
                    P.h_leave[n_pivots] = leave;
                }
                if (s_local >= 0) P.h_col[n_pivots] = (int32_t)(lo + s_local);  // exactly one CTA owns the column
            }
        }
        ++n_pivots;
        pending = true;
        pr = r;
        ps = s_local;
        __syncthreads();
    }

    // ---- write the slice, the labels and the state back ----
    __syncthreads();
    for (int64_t e = tid; e < R * (wl + 1); e += ONCHIP_THREADS) {
        const int64_t i = e / (wl + 1);
        const int j = (int)(e - i * (wl + 1));
        if (j < wl) P.T[i * P.ld + lo + j] = Tl[i * stride + j];
        else if (g == G - 1) P.T[i * P.ld + C - 1] = Tl[i * stride + j];
    }
    for (int j = tid; j < wl; j += ONCHIP_THREADS) P.collab[lo + j] = cl[j];
    if (g == 0) {
        for (int64_t i = tid; i < R; i += ONCHIP_THREADS) P.rowlab[i] = rl[i];
        if (tid == 0) {
            P.st->done = 1;
            P.st->status = status;
            P.st->have_pivot = 0;
            P.st->pend = 0;
            P.st->n_pivots = n_pivots;
        }
    }
}

}  // namespace b200lp
