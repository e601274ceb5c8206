// kernels_batched.cuh -- B independent small LPs, one warp per LP, the whole two-phase solve in one launch.
//
// The condensed tableau ((m+2) x (n + #>= rows + 1) doubles; 9 KB for the 20 x 30 LPs of BASELINE config 3)
// lives in shared memory for the life of the LP, so HBM sees only the inputs (A, b, c, ops) and the outputs.
// Each lane owns whole columns (j = lane, lane+32, ...): the rank-1 update needs no intra-warp exchange beyond
// a broadcast read of the saved pivot column, and consecutive lanes touch consecutive doubles (conflict-free).
// The row stride is odd so that the strided column reads of the ratio test are conflict-free too.
// Control flow is warp-uniform; every decision is a warp-shuffle reduction with a total order, so the pivot
// sequence is bit-identical to oracle/simplex_oracle.c.
//
// Reference seam: one SolverController.run() per problem
// (/root/reference/app/controllers/ui_controller.py:194-195, solver_controller.py:53-120).
#pragma once
#include "common.cuh"

namespace b200lp {

struct BatchedParams {
    int64_t B;
    int32_t m, n;
    int32_t ld;         // odd row stride of the shared-memory tableau, >= n + m + 1
    int32_t rule;
    int32_t max_pivots;
    int32_t auto_budget;  // 1: budget = 200 * (m + C) + 10000 per LP, with the Dantzig -> Bland continuation
    int32_t log_cap;
    double eps_cost, eps_pivot, eps_feas;
    const double* A;
    const double* b;
    const double* c;
    const int8_t* ops;
    int32_t* status;
    double* fun;
    double* x;
    int32_t* n_pivots;
    int32_t* piv_log;
    size_t warp_bytes;  // shared memory per warp
    unsigned int* next;  // work counter of this launch (zeroed by the host)
};

struct WarpLP {
    double* T;
    double* colbuf;
    int32_t* rowlab;
    int32_t* collab;
    int32_t m, n, R, C, ld, art_base;
};

// IEEE division whose numerator may be an exact zero (slack / surplus entries, degenerate right-hand sides): a zero
// numerator sends div.rn.f64 down its slow path (a ~35-instruction subroutine the whole warp waits for; measured at two
// calls per pivot on the config-3 LPs), so the quotient of a zero is formed as num * den instead -- the same +-0 for
// every finite non-zero den -- and the divider only ever sees a non-zero numerator.
__device__ __forceinline__ double div_maybe_zero(double num, double den) {
    const bool z = (num == 0.0);
    double safe = z ? 1.0 : num;
    asm volatile("" : "+d"(safe));  // opaque: otherwise the compiler divides `num` itself and selects afterwards
    const double q = safe / den;
    return z ? num * den : q;
}

// The rank-1 update of one lane's columns j (and j + 32 when TWO) over ALL rows, without per-element predicates: row r
// gets a throw-away value and is overwritten with q right after; column s was zeroed when it was saved, so what the
// chain loads there is the 0 of the arithmetic contract fma(-col_i, 1/p, 0).  GROUP rows are loaded together before
// their stores, so the shared-memory latency of one row hides behind the others (a store cannot move above an earlier
// load of unknown alias); with TWO the broadcast load of col_i, the address arithmetic and the loop control serve two
// elements.  has1: this lane really owns a column j + 32 (lanes beyond C only skip their accesses).
template <int GROUP, bool TWO>
__device__ __forceinline__ void wlp_update_cols(const WarpLP& w, int j, double q0, double q1, bool has1) {
    double* cell = w.T + j;
    const double* cb = w.colbuf;
    const int ld = w.ld;
    int i = 0;
    for (; i + GROUP <= w.R; i += GROUP) {
        double cc[GROUP], t0[GROUP], t1[GROUP];
#pragma unroll
        for (int u = 0; u < GROUP; ++u) {
            cc[u] = cb[u];
            t0[u] = cell[u * ld];
            if (TWO) t1[u] = has1 ? cell[u * ld + 32] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < GROUP; ++u) {
            cell[u * ld] = __fma_rn(-cc[u], q0, t0[u]);
            if (TWO && has1) cell[u * ld + 32] = __fma_rn(-cc[u], q1, t1[u]);
        }
        cell += GROUP * ld;
        cb += GROUP;
    }
    for (; i < w.R; ++i) {
        const double c0 = cb[0];
        const double t0 = cell[0];
        double t1 = 0.0;
        if (TWO && has1) t1 = cell[32];
        cell[0] = __fma_rn(-c0, q0, t0);
        if (TWO && has1) cell[32] = __fma_rn(-c0, q1, t1);
        cell += ld;
        ++cb;
    }
}

// GROUP: rows whose loads are issued together before their stores (1 = plain loop, fewer registers: better for
// short tableaux where occupancy matters more; 4 = measured best for the 22-row tableaux of config 3).
template <int GROUP>
__device__ __forceinline__ void wlp_pivot(const WarpLP& w, int r, int s, int lane) {
    // save column s and clear it in place (see wlp_update_cols)
    for (int i = lane; i < w.R; i += 32) {
        double* e = w.T + i * w.ld + s;
        w.colbuf[i] = *e;
        *e = 0.0;
    }
    __syncwarp();
    const double p = w.colbuf[r];
    const double inv_p = 1.0 / p;
    double* rowr = w.T + r * w.ld;
    if (w.C > 32) {
        for (int j = lane; j < w.C; j += 64) {
            const bool has1 = j + 32 < w.C;
            const double q0 = (j == s) ? inv_p : div_maybe_zero(rowr[j], p);
            const double q1 = (j + 32 == s) ? inv_p : div_maybe_zero(has1 ? rowr[j + 32] : 1.0, p);
            wlp_update_cols<GROUP, true>(w, j, q0, q1, has1);
            rowr[j] = q0;
            if (has1) rowr[j + 32] = q1;
        }
    } else {
        if (lane < w.C) {
            const double q0 = (lane == s) ? inv_p : div_maybe_zero(rowr[lane], p);
            wlp_update_cols<GROUP, false>(w, lane, q0, 0.0, false);
            rowr[lane] = q0;
        }
    }
    __syncwarp();
    if (lane == 0) {
        const int32_t leave = w.rowlab[r];
        w.rowlab[r] = w.collab[s];
        w.collab[s] = leave;
    }
    __syncwarp();
}

template <bool BLAND>
__device__ __forceinline__ int wlp_price(const WarpLP& w, int obj_row, double eps_cost, int lane) {
    Key k = key_none();
    const double* d = w.T + obj_row * w.ld;
    for (int j = lane; j < w.C - 1; j += 32) {
        const int32_t lab = w.collab[j];
        const double v = d[j];
        if (lab < w.art_base && v < -eps_cost) {
            Key c;
            c.v = v;
            c.lab = lab;
            c.pos = j;
            k = key_min<BLAND>(k, c);
        }
    }
    k = warp_key_min<BLAND>(k);
    return k.pos;
}

__device__ __forceinline__ int wlp_ratio(const WarpLP& w, int s, double eps_pivot, int lane) {
    Key k = key_none();
    for (int i = lane; i < w.m; i += 32) {
        const int32_t lab = w.rowlab[i];
        const double a = w.T[i * w.ld + s];
        if (lab >= 0 && a > eps_pivot) {
            Key c;
            c.v = div_maybe_zero(w.T[i * w.ld + w.C - 1], a);
            c.lab = lab;
            c.pos = i;
            k = key_min<false>(k, c);
        }
    }
    k = warp_key_min<false>(k);
    return k.pos;
}

struct WarpRun {
    int32_t n_pivots;
    int32_t max_pivots;
    int32_t log_cap;
    int32_t* log;  // this LP's [log_cap][2] slice or nullptr
};

__device__ __forceinline__ void wlp_log(WarpRun& run, int r, int s, int lane) {
    if (lane == 0 && run.log && run.n_pivots < run.log_cap) {
        run.log[2 * run.n_pivots + 0] = r;
        run.log[2 * run.n_pivots + 1] = s;
    }
    run.n_pivots++;
}

// one phase on objective row obj_row; returns a B200LP_STATUS_* code (0 = optimal for this row)
template <int GROUP>
__device__ __forceinline__ int wlp_run_phase(const WarpLP& w, int obj_row, int rule, double eps_cost, double eps_pivot,
                                             WarpRun& run, int lane) {
    for (;;) {
        if (run.n_pivots >= run.max_pivots) return 1;
        const int s = rule == 1 ? wlp_price<true>(w, obj_row, eps_cost, lane) : wlp_price<false>(w, obj_row, eps_cost, lane);
        if (s < 0) return 0;
        const int r = wlp_ratio(w, s, eps_pivot, lane);
        if (r < 0) return 3;
        wlp_log(run, r, s, lane);
        wlp_pivot<GROUP>(w, r, s, lane);
    }
}

template <int GROUP>
__device__ __forceinline__ int wlp_drive_out(const WarpLP& w, double eps_pivot, WarpRun& run, int lane) {
    for (int i = 0; i < w.m; ++i) {
        if (w.rowlab[i] < w.art_base) continue;
        Key k = key_none();
        const double* row = w.T + i * w.ld;
        for (int j = lane; j < w.C - 1; j += 32) {
            const int32_t lab = w.collab[j];
            const double a = fabs(row[j]);
            if (lab < w.art_base && a > eps_pivot) {
                Key c;
                c.v = -a;
                c.lab = lab;
                c.pos = j;
                k = key_min<false>(k, c);
            }
        }
        k = warp_key_min<false>(k);
        if (k.pos < 0) {
            __syncwarp();
            if (lane == 0) w.rowlab[i] = -1 - w.rowlab[i];
            __syncwarp();
            continue;
        }
        if (run.n_pivots >= run.max_pivots) return 1;
        wlp_log(run, i, k.pos, lane);
        wlp_pivot<GROUP>(w, i, k.pos, lane);
    }
    return 0;
}

// the whole two-phase solve of LP `lp` by one warp in its slice `base` of shared memory
template <int GROUP>
__device__ __forceinline__ void wlp_solve_one(const BatchedParams& P, const int64_t lp, uint8_t* base, const int lane) {
    const int m = P.m, n = P.n, ld = P.ld, R = m + 2;

    WarpLP w;
    w.T = reinterpret_cast<double*>(base);
    w.colbuf = w.T + (size_t)R * ld;
    w.rowlab = reinterpret_cast<int32_t*>(w.colbuf + R);
    w.collab = w.rowlab + R;
    w.m = m;
    w.n = n;
    w.ld = ld;
    w.art_base = n + m;

    const double* A = P.A + lp * m * n;
    const double* b = P.b + lp * m;
    const double* c = P.c + lp * n;
    const int8_t* ops = P.ops + lp * m;

    // ---- normalise rows (b < 0 -> negate), place surplus columns, label the slack / artificial basis ----
    for (int e = lane; e < R * ld; e += 32) w.T[e] = 0.0;
    for (int j = lane; j < ld; j += 32) w.collab[j] = j < n ? j : -1;
    __syncwarp();
    int n_ge = 0, n_art = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        int op = 0;
        bool neg = false;
        if (i < m) {
            op = ops[i];
            neg = b[i] < 0.0;
            if (neg && op != 2) op = (op == 0) ? 1 : 0;
        }
        const unsigned ge_mask = __ballot_sync(0xffffffffu, i < m && op == 1);
        const unsigned art_mask = __ballot_sync(0xffffffffu, i < m && op != 0);
        if (i < m) {
            const double bi = b[i];
            w.T[i * ld + (ld - 1)] = neg ? -bi : bi;  // parked in the last slot; moved to column C-1 below
            if (op == 1) {
                const int k = n_ge + __popc(ge_mask & ((1u << lane) - 1u));
                w.T[i * ld + n + k] = -1.0;
                w.collab[n + k] = n + i;
            }
            w.rowlab[i] = (op == 0) ? n + i : w.art_base + i;
            // remember the sign flip in colbuf for the coefficient load below
            w.colbuf[i] = neg ? -1.0 : 1.0;
        }
        n_ge += __popc(ge_mask);
        n_art += __popc(art_mask);
    }
    __syncwarp();
    const int C = n + n_ge + 1;
    const int n_obj = n_art > 0 ? 2 : 1;
    w.C = C;
    w.R = m + n_obj;
    if (C - 1 != ld - 1) {
        for (int i = lane; i < m; i += 32) {
            const double v = w.T[i * ld + (ld - 1)];
            w.T[i * ld + (ld - 1)] = 0.0;
            w.T[i * ld + (C - 1)] = v;
        }
    }
    if (lane == 0) {
        w.collab[C - 1] = -1;
        w.rowlab[m] = -1;
        w.rowlab[m + 1] = -1;
    }
    __syncwarp();
    for (int i = 0; i < m; ++i) {
        const bool flip = w.colbuf[i] < 0.0;  // warp-uniform
        for (int j = lane; j < n; j += 32) {
            const double a = A[i * n + j];
            w.T[i * ld + j] = flip ? -a : a;
        }
    }
    for (int j = lane; j < n; j += 32) w.T[m * ld + j] = c[j];
    __syncwarp();
    if (n_obj == 2) {
        for (int j = lane; j < C; j += 32) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i)
                if (w.rowlab[i] >= w.art_base) acc = __dadd_rn(acc, w.T[i * ld + j]);
            w.T[(m + 1) * ld + j] = -acc;
        }
    }
    __syncwarp();

    // ---- two-phase solve ----
    WarpRun run;
    run.n_pivots = 0;
    run.max_pivots = P.max_pivots;
    run.log_cap = P.log_cap;
    run.log = P.piv_log ? P.piv_log + lp * (int64_t)P.log_cap * 2 : nullptr;
    if (run.log)
        for (int e = lane; e < 2 * P.log_cap; e += 32) run.log[e] = -1;
    __syncwarp();
    int st = 0;
    int rule = P.rule;
    const int cap = P.auto_budget ? 200 * (m + C) + 10000 : P.max_pivots;
    run.max_pivots = cap;
    if (n_obj == 2) {
        st = wlp_run_phase<GROUP>(w, m + 1, rule, P.eps_cost, P.eps_pivot, run, lane);
        if (st == 1 && P.auto_budget && rule == 0) {  // Dantzig stalled: continue under Bland (cannot cycle)
            rule = 1;
            run.max_pivots += cap;
            st = wlp_run_phase<GROUP>(w, m + 1, rule, P.eps_cost, P.eps_pivot, run, lane);
        }
        if (st == 3) st = 4;
        if (st == 0 && w.T[(m + 1) * ld + C - 1] < -P.eps_feas) st = 2;
        if (st == 0) st = wlp_drive_out<GROUP>(w, P.eps_pivot, run, lane);
    }
    if (st == 0) {
        st = wlp_run_phase<GROUP>(w, m, rule, P.eps_cost, P.eps_pivot, run, lane);
        if (st == 1 && P.auto_budget && rule == 0) {
            rule = 1;
            run.max_pivots += cap;
            st = wlp_run_phase<GROUP>(w, m, rule, P.eps_cost, P.eps_pivot, run, lane);
        }
    }

    // ---- results ----
    if (P.x) {
        double* x = P.x + lp * n;
        for (int j = lane; j < n; j += 32) x[j] = 0.0;
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            const int32_t lab = w.rowlab[i];
            if (lab >= 0 && lab < n) x[lab] = w.T[i * ld + C - 1];
        }
    }
    if (lane == 0) {
        P.status[lp] = st;
        P.fun[lp] = -w.T[m * ld + C - 1];
        P.n_pivots[lp] = run.n_pivots;
    }
}

// Persistent warps: the grid is sized to the resident capacity of the chip and every warp draws its next LP from a
// device counter, so a warp whose LP ends early (infeasible after two pivots, 10 instead of 60 pivots) starts another at
// once instead of idling until the slowest LP of its CTA has finished.
template <int GROUP>
__global__ void __launch_bounds__(256) k_solve_batched(const BatchedParams P) {
    extern __shared__ __align__(16) uint8_t smem_batched[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* base = smem_batched + (size_t)warp * P.warp_bytes;
    for (;;) {
        unsigned int v = 0;
        if (lane == 0) v = atomicAdd(P.next, 1u);
        v = __shfl_sync(0xffffffffu, v, 0);
        if ((int64_t)v >= P.B) return;
        wlp_solve_one<GROUP>(P, (int64_t)v, base, lane);
        __syncwarp();  // the slice is reused by the next LP
    }
}

}  // namespace b200lp
