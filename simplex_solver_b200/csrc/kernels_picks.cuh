// kernels_picks.cuh -- all K decisions of a look-ahead block in ONE persistent cooperative kernel over every SM.
//
// k_pick_cluster<.., BLOCKED> takes one pick per launch on one 16-CTA cluster: with t pivots pending it re-reads the
// whole history, t * (R + C) * 8 bytes, through the path of ONE GPC to L2 (~0.6 TB/s): 12.5 us + 0.47 us per pending
// step.  Here CTA g owns a slice of columns and a slice of rows for the whole block and keeps THEIR history -- q_u of its
// columns, col_u of its rows -- in shared memory, so the replay of a pick reads nothing from L2 but the 2 t scalars
// col_u[r] and q_u[s].  Per pick: [row part of the previous pivot + pricing on the own columns] -> grid barrier -> every
// CTA reduces the G candidates (same total order, same winner) -> [replay of the entering column on the own rows + ratio
// test] -> grid barrier -> every CTA reduces the G candidates.  Labels, pending steps and the pivot count are replicated
// in every CTA (all take identical decisions); CTA 0 writes the pivot history and, at the end, the DevState.
// The row part of the block's last pivot is done before the kernel ends (no k_blk_row launch).
// Arithmetic, masks and tie-breaking are those of k_pick_cluster: pivots and buffers are bit-identical.
#pragma once
#include "kernels_cluster.cuh"
#include "kernels_onchip.cuh"

namespace b200lp {

constexpr int PK_THREADS = 256;
constexpr size_t PK_SMEM_MAX = 200 * 1024;  // history of the own columns and rows, K steps deep

struct PickPartB {  // ratio-test candidate of one CTA with what the others need from its row
    Key k;          // (ratio, basic variable id, row)
    double a;       // col_t[row] = the pivot element if this candidate wins
    double rhs;     // current right-hand side of the row
};

struct PicksArgs {
    PickArgs A;
    int32_t K;       // picks to take in this launch (<= BLK_KMAX)
    int32_t wC, wR;  // columns / rows per CTA
    unsigned long long* barrier;  // monotonic arrival counter, zeroed before the launch
    Key* partA;        // [G]
    PickPartB* partB;  // [G]
    int32_t* error;    // set when a grid barrier times out (a CTA went missing): the host reports it
};

// grid barrier with a watchdog: a barrier that cannot complete must not hang the GPU
__device__ __forceinline__ bool grid_barrier_wd(unsigned long long* counter, unsigned long long target) {
    __shared__ int ok_s;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
        unsigned long long v;
        unsigned int spins = 0;
        int ok = 1;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
            // polls are counted (a round trip to L2 each, ~0.4 us) rather than timed: a clock read between two polls
            // delays the exit from the barrier (as in k_solve_onchip)
            if (v < target && ++spins > (1u << 23)) {  // a few seconds
                ok = 0;
                break;
            }
        } while (v < target);
        ok_s = ok;
    }
    __syncthreads();
    return ok_s != 0;
}

template <bool BLAND>
__global__ void __launch_bounds__(PK_THREADS, 1) k_blk_picks(const PicksArgs P) {
    extern __shared__ __align__(16) double dyn_pk[];
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sx[BLK_KMAX];
    __shared__ Key skA[PK_THREADS / 32], skW[PK_THREADS / 32], skB[PK_THREADS / 32], skV[PK_THREADS / 32];  // one per reduction
    __shared__ PickPartB bcB;
    const PickArgs& A = P.A;
    DevState* st = A.st;
    if (st->done) return;  // grid-uniform
    const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x;
    const int64_t R = A.R, C = A.C, ld = A.ld, m = A.m;
    const int wC = P.wC, wR = P.wR;
    const int64_t j0 = (int64_t)g * wC, j1 = min(C, j0 + wC);
    const int64_t i0 = (int64_t)g * wR, i1 = min(R, i0 + wR);
    double* qh = dyn_pk;                                   // [K][wC]  q_u of the own columns
    double* ch = qh + (size_t)P.K * wC;                    // [K][wR]  col_u of the own rows
    int32_t* cl = reinterpret_cast<int32_t*>(ch + (size_t)P.K * wR);  // labels of the own columns
    int32_t* rl = cl + wC;                                 // labels of the own rows (-1: not a constraint row)
    for (int64_t j = j0 + tid; j < j1; j += PK_THREADS) cl[j - j0] = A.collab[j];
    for (int64_t i = i0 + tid; i < i1; i += PK_THREADS) rl[i - i0] = i < m ? A.rowlab[i] : -1;

    long long n_piv = st->n_pivots;
    const long long max_pivots = st->max_pivots;
    if (n_piv != A.B.pend->base || st->pend) {  // grid-uniform: the launch must start right after a flush
        if (g == 0 && tid == 0) {
            *P.error = 2;
            st->done = 1;
            st->status = 4;
            st->have_pivot = 0;
        }
        return;
    }
    int t = 0;             // pivots pending in this block (the launch starts right after a flush)
    bool pend = false;     // the row part of pivot t - 1 is due
    int r_prev = -1, s_prev = -1, enter_prev = -1, leave_prev = -1;
    double p_prev = 0.0, inv_prev = 0.0, q_rhs_prev = 0.0, best_prev = 0.0;
    int status = -1;       // >= 0: the loop has ended
    bool no_candidate = false;
    bool ok = true;
    unsigned long long round = 0;
    __syncthreads();

    // row part of pivot t - 1 on the own columns (q_{t-1}, current objective row) and own rows (current RHS); returns
    // this thread's pricing candidate
    auto row_part = [&]() {
        Key k = key_none();
        const int tp = t - 1;
        if (tid < tp) sx[tid] = __ldcg(A.B.colP + (int64_t)tid * A.B.Rpad + r_prev);  // col_u[r], u < tp
        const double c_obj = __ldcg(A.B.colP + (int64_t)tp * A.B.Rpad + A.obj_row);
        __syncthreads();
        for (int64_t j = j0 + tid; j < j1; j += PK_THREADS) {
            double v = A.T[(int64_t)r_prev * ld + j];
            for (int u = 0; u < tp; ++u)
                v = blk_step(v, r_prev == sr[u], j == ss[u], sx[u], qh[(size_t)u * wC + (j - j0)], sinv[u]);
            const double q = (j == s_prev) ? inv_prev : v / p_prev;
            qh[(size_t)tp * wC + (j - j0)] = q;
            A.B.qP[(int64_t)tp * A.B.Cpad + j] = q;
            const double d = blk_step(A.B.objcur[j], false, j == s_prev, c_obj, q, inv_prev);
            A.B.objcur[j] = d;
            if (j < C - 1) {
                const int32_t lab = cl[j - j0];
                if (lab < A.art_base && d < -A.eps_cost) {
                    Key c;
                    c.v = d;
                    c.lab = lab;
                    c.pos = (int32_t)j;
                    k = key_min<BLAND>(k, c);
                }
            }
        }
        for (int64_t i = i0 + tid; i < i1; i += PK_THREADS)
            A.B.rhscur[i] = (i == r_prev) ? q_rhs_prev : __fma_rn(-ch[(size_t)tp * wR + (i - i0)], q_rhs_prev, A.B.rhscur[i]);
        return k;
    };

    for (int it = 0; it < P.K; ++it) {
        // ------------------------------------------------ columns: entering variable ---------------------------------
        Key k = key_none();
        if (pend) {
            k = row_part();
        } else {
            for (int64_t j = j0 + tid; j < min(j1, C - 1); j += PK_THREADS) {
                const int32_t lab = cl[j - j0];
                const double d = A.B.objcur[j];
                if (lab < A.art_base && d < -A.eps_cost) {
                    Key c;
                    c.v = d;
                    c.lab = lab;
                    c.pos = (int32_t)j;
                    k = key_min<BLAND>(k, c);
                }
            }
        }
        pend = false;
        k = block_key_min_all<BLAND>(k, skA);
        if (tid == 0) P.partA[g] = k;
        ok = grid_barrier_wd(P.barrier, ++round * (unsigned long long)G);
        if (!ok) break;
        Key w = key_none();
        for (int x = tid; x < G; x += PK_THREADS) {
            Key c;
            const double2 raw = __ldcg(reinterpret_cast<const double2*>(P.partA + x));
            c.v = raw.x;
            c.lab = (int32_t)(__double_as_longlong(raw.y) & 0xffffffffll);
            c.pos = (int32_t)(__double_as_longlong(raw.y) >> 32);
            w = key_min<BLAND>(w, c);
        }
        const Key win = block_key_min_all<BLAND>(w, skW);
        if (n_piv >= max_pivots || win.lab == B200LP_NO_LAB) {
            status = n_piv >= max_pivots ? 1 : 0;  // LIMIT is checked first, as in the oracle
            no_candidate = win.lab == B200LP_NO_LAB;
            break;
        }
        const int s = win.pos;

        // ------------------------------------------------ rows: leaving variable -------------------------------------
        if (tid < t) sx[tid] = __ldcg(A.B.qP + (int64_t)tid * A.B.Cpad + s);  // q_u[s]; q_{t-1} was written before the barrier
        __syncthreads();
        Key kb = key_none();
        double ka = 0.0, krhs = 0.0;
        for (int64_t i = i0 + tid; i < i1; i += PK_THREADS) {
            double a = A.T[i * ld + s];
            for (int u = 0; u < t; ++u)
                a = blk_step(a, i == sr[u], s == ss[u], ch[(size_t)u * wR + (i - i0)], sx[u], sinv[u]);
            ch[(size_t)t * wR + (i - i0)] = a;
            A.B.colP[(int64_t)t * A.B.Rpad + i] = a;
            const int32_t lab = rl[i - i0];
            if (lab >= 0 && a > A.eps_pivot) {
                const double rhs = A.B.rhscur[i];
                Key c;
                c.v = rhs / a;
                c.lab = lab;
                c.pos = (int32_t)i;
                if (key_less_val(c, kb)) {
                    kb = c;
                    ka = a;
                    krhs = rhs;
                }
            }
        }
        {
            const Key wb = block_key_min_all<false>(kb, skB);
            if (tid == 0) bcB.k = wb;
            if (wb.lab != B200LP_NO_LAB && kb.lab != B200LP_NO_LAB && kb.pos == wb.pos) {  // the one thread that holds it
                bcB.a = ka;
                bcB.rhs = krhs;
            }
            __syncthreads();
            if (tid == 0) P.partB[g] = bcB;
        }
        ok = grid_barrier_wd(P.barrier, ++round * (unsigned long long)G);
        if (!ok) break;
        kb = key_none();
        for (int x = tid; x < G; x += PK_THREADS) {
            const double2 raw = __ldcg(reinterpret_cast<const double2*>(&P.partB[x].k));
            Key c;
            c.v = raw.x;
            c.lab = (int32_t)(__double_as_longlong(raw.y) & 0xffffffffll);
            c.pos = (int32_t)(__double_as_longlong(raw.y) >> 32);
            if (key_less_val(c, kb)) {
                kb = c;
                ka = __ldcg(&P.partB[x].a);
                krhs = __ldcg(&P.partB[x].rhs);
            }
        }
        {
            const Key wb = block_key_min_all<false>(kb, skV);
            if (tid == 0) bcB.k = wb;
            if (wb.lab != B200LP_NO_LAB && kb.lab != B200LP_NO_LAB && kb.pos == wb.pos) {
                bcB.a = ka;
                bcB.rhs = krhs;
            }
            __syncthreads();
        }
        if (bcB.k.lab == B200LP_NO_LAB) {
            status = 3;  // UNBOUNDED
            break;
        }
        const int r = bcB.k.pos;
        const int32_t leave = bcB.k.lab;
        const double p = bcB.a;
        const double inv_p = 1.0 / p;
        const double q_rhs = bcB.rhs / p;
        if (tid == 0) {
            sr[t] = r;
            ss[t] = s;
            sinv[t] = inv_p;
            if (r >= i0 && r < i1) {
                rl[r - i0] = win.lab;
                A.rowlab[r] = win.lab;
            }
            if (s >= j0 && s < j1) {
                cl[s - j0] = leave;
                A.collab[s] = leave;
            }
            if (g == 0 && n_piv < A.hist_cap) {
                A.h_row[n_piv] = r;
                A.h_col[n_piv] = s;
                A.h_enter[n_piv] = win.lab;
                A.h_leave[n_piv] = leave;
            }
        }
        r_prev = r;
        s_prev = s;
        p_prev = p;
        inv_prev = inv_p;
        q_rhs_prev = q_rhs;
        enter_prev = win.lab;
        leave_prev = leave;
        best_prev = win.v;
        pend = true;
        ++t;
        ++n_piv;
        __syncthreads();
    }
    if (ok && pend) row_part();  // the block's last pivot: its row part is due before the flush

    if (g == 0 && tid == 0) {
        if (!ok) {
            *P.error = 1;
            st->done = 1;
            st->status = 4;
            st->have_pivot = 0;
        } else {
            for (int u = 0; u < t; ++u) {
                A.B.pend->r[u] = sr[u];
                A.B.pend->s[u] = ss[u];
                A.B.pend->inv_p[u] = sinv[u];
            }
            A.B.pend->q_rhs = q_rhs_prev;
            st->n_pivots = n_piv;
            st->pend = 0;
            if (t > 0) {  // the last pivot taken, as k_pick_cluster leaves it
                st->r = r_prev;
                st->p = p_prev;
                st->inv_p = inv_prev;
                st->leave_lab = leave_prev;
                st->s = s_prev;
                st->enter_lab = enter_prev;
                st->best_val = best_prev;
            }
            st->have_pivot = status < 0;
            if (status >= 0) {
                st->done = 1;
                st->status = status;
                if (no_candidate) {
                    st->s = -1;
                    st->enter_lab = -1;
                }
            }
        }
    }
}

}  // namespace b200lp
