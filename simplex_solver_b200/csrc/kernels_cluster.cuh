// kernels_cluster.cuh -- both decisions of a pivot (entering column, leaving row) in ONE launch on ONE thread-block
// cluster: the two grid-wide reductions are finished through distributed shared memory and the hardware cluster barrier
// instead of "partials in global memory + fence + atomic ticket + last CTA re-reads".  Replaces k_price + k_ratio in the
// rank-1 loop and k_blk_rowprice + k_blk_ratio in the look-ahead loop (non-sharded tableaux).
//
//   phase A (columns): [rank-1: finish the previous pivot's row scaling | look-ahead: row part of the previous pivot]
//                      + argmin over the objective row  -> cluster reduction #1 -> entering column s
//   phase B (rows)   : entering column gathered (look-ahead: + replay of the pending steps) and saved, masked min-ratio
//                      -> cluster reduction #2 -> leaving row r; block 0 does the bookkeeping.
//
// Data written in phase A by one CTA and read in phase B by another goes through L2 (ld.global.cg), ordered by the
// cluster barrier.  Decisions use the total orders of common.cuh, so the result is independent of the cluster size.
#pragma once
#include <cooperative_groups.h>

#include "kernels_blocked.cuh"

namespace b200lp {

namespace cg = cooperative_groups;

constexpr int CL_THREADS = 1024;
// Replay operands loaded 8 at a time ahead of the chain (blk_replay<.., true>): measured SLOWER here (45.3 vs 43.5 us per
// pivot at K = 32).  The cluster sits in one GPC and its replay traffic, t * (R + C) * 8 bytes per pick, moves at the
// ~0.6 TB/s of that GPC's path to L2 however many loads are in flight.
#ifndef PICK_BATCH
#define PICK_BATCH false
#endif

struct PickArgs {
    double* T;
    int64_t R, m, C, ld, obj_row;
    int32_t* rowlab;
    int32_t* collab;
    int32_t art_base;
    double eps_cost, eps_pivot;
    DevState* st;
    double* col;  // rank-1 loop: contiguous copy of the entering column for the update kernel
    BlkBuffers B; // look-ahead loop
    int32_t* h_row;
    int32_t* h_col;
    int32_t* h_enter;
    int32_t* h_leave;
    int64_t hist_cap;
};

// min over the cluster of one Key per CTA: every CTA ends up with the same winner
template <bool BY_LABEL>
__device__ __forceinline__ Key cluster_key_min(cg::cluster_group& cluster, Key mine, Key* slot, Key* sk, Key* bc) {
    mine = block_key_min<BY_LABEL>(mine, sk);
    if (threadIdx.x == 0) *slot = mine;
    cluster.sync();
    if (threadIdx.x < 32) {
        Key k = key_none();
        if (threadIdx.x < cluster.num_blocks()) k = *cluster.map_shared_rank(slot, threadIdx.x);
        k = warp_key_min<BY_LABEL>(k);
        if (threadIdx.x == 0) *bc = k;
    }
    __syncthreads();
    return *bc;
}

template <bool BLAND, bool BLOCKED>
__global__ void __launch_bounds__(CL_THREADS, 1) k_pick_cluster(const PickArgs A) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ Key sk[CL_THREADS / 32];
    __shared__ Key slot_price, slot_ratio, bc;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sx[BLK_KMAX];
    DevState* st = A.st;
    if (st->done) return;  // cluster-uniform: nobody writes the state before the last barrier
    const int64_t gtid = (int64_t)cluster.block_rank() * CL_THREADS + threadIdx.x;
    const int64_t nthr = (int64_t)cluster.num_blocks() * CL_THREADS;
    const int64_t R = A.R, C = A.C, ld = A.ld;
    const long long n_piv = st->n_pivots;
    const long long base = BLOCKED ? A.B.pend->base : 0;

    // ------------------------------------------------ phase A: columns ------------------------------------------------
    Key k = key_none();
    if (BLOCKED) {
        const bool row_due = st->pend != 0;
        if (row_due) {
            const int t = (int)(n_piv - 1 - base);
            const int r = st->r, s = st->s;
            const double p = st->p, inv_p = st->inv_p;
            if (threadIdx.x < t) {
                sr[threadIdx.x] = A.B.pend->r[threadIdx.x];
                ss[threadIdx.x] = A.B.pend->s[threadIdx.x];
                sinv[threadIdx.x] = A.B.pend->inv_p[threadIdx.x];
                sx[threadIdx.x] = A.B.colP[(int64_t)threadIdx.x * A.B.Rpad + r];  // col_u[r]
            }
            __syncthreads();
            const double* colT = A.B.colP + (int64_t)t * A.B.Rpad;
            double* qT = A.B.qP + (int64_t)t * A.B.Cpad;
            const double c_obj = colT[A.obj_row];
            const double q_rhs = A.B.pend->q_rhs;
            for (int64_t j = gtid; j < C; j += nthr) {
                double v = A.T[(int64_t)r * ld + j];
                v = blk_replay<true, PICK_BATCH>(v, t, A.B.qP + j, A.B.Cpad, sx, sr, ss, sinv, r, j);
                const double q = (j == s) ? inv_p : v / p;
                qT[j] = q;
                const double d = blk_step(A.B.objcur[j], false, j == s, c_obj, q, inv_p);
                A.B.objcur[j] = d;
                if (j < C - 1) {
                    const int32_t lab = A.collab[j];
                    if (lab < A.art_base && d < -A.eps_cost) {
                        Key c;
                        c.v = d;
                        c.lab = lab;
                        c.pos = (int32_t)j;
                        k = key_min<BLAND>(k, c);
                    }
                }
            }
            for (int64_t i = gtid; i < R; i += nthr)
                A.B.rhscur[i] = (i == r) ? q_rhs : __fma_rn(-colT[i], q_rhs, A.B.rhscur[i]);
            __syncthreads();  // sr/ss/sinv/sx are reloaded in phase B
        } else {
            for (int64_t j = gtid; j < C - 1; j += nthr) {
                const int32_t lab = A.collab[j];
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
    } else {
        if (st->pend) {
            const int r = st->r, s = st->s;
            const double p = st->p, inv_p = st->inv_p;
            double* row = A.T + (int64_t)r * ld;
            for (int64_t j = gtid; j < C; j += nthr) {
                const double v = row[j];
                row[j] = (j == s) ? inv_p : v / p;
            }
        }
        const double* d = A.T + A.obj_row * ld;
        for (int64_t j = gtid; j < C - 1; j += nthr) {
            const int32_t lab = A.collab[j];
            const double v = d[j];
            if (lab < A.art_base && v < -A.eps_cost) {
                Key c;
                c.v = v;
                c.lab = lab;
                c.pos = (int32_t)j;
                k = key_min<BLAND>(k, c);
            }
        }
    }
    const Key win = cluster_key_min<BLAND>(cluster, k, &slot_price, sk, &bc);
    const bool leader = cluster.block_rank() == 0 && threadIdx.x == 0;
    if (n_piv >= st->max_pivots || win.lab == B200LP_NO_LAB) {
        cluster.sync();  // nobody leaves while its shared memory may still be read; all state reads are done
        if (leader) {
            st->pend = 0;
            st->have_pivot = 0;
            st->done = 1;
            st->status = (n_piv >= st->max_pivots) ? 1 : 0;  // LIMIT is checked first, as in the oracle
            if (win.lab == B200LP_NO_LAB) {
                st->s = -1;
                st->enter_lab = -1;
            }
        }
        return;
    }
    const int s = win.pos;

    // ------------------------------------------------ phase B: rows ---------------------------------------------------
    int t = 0;
    double* colT = A.col;
    if (BLOCKED) {
        t = (int)(n_piv - base);
        colT = A.B.colP + (int64_t)t * A.B.Rpad;
        if (threadIdx.x < t) {
            sr[threadIdx.x] = A.B.pend->r[threadIdx.x];
            ss[threadIdx.x] = A.B.pend->s[threadIdx.x];
            sinv[threadIdx.x] = A.B.pend->inv_p[threadIdx.x];
            sx[threadIdx.x] = __ldcg(A.B.qP + (int64_t)threadIdx.x * A.B.Cpad + s);  // q_u[s]; q_{t-1} is fresh from phase A
        }
        __syncthreads();
    }
    k = key_none();
    for (int64_t i = gtid; i < R; i += nthr) {
        double a = __ldcg(A.T + i * ld + s);
        double rhs;
        if (BLOCKED) {
            a = blk_replay<false, PICK_BATCH>(a, t, A.B.colP + i, A.B.Rpad, sx, sr, ss, sinv, i, s);
            rhs = __ldcg(A.B.rhscur + i);
        } else {
            rhs = __ldcg(A.T + i * ld + C - 1);
        }
        colT[i] = a;
        if (i < A.m) {
            const int32_t lab = A.rowlab[i];
            if (lab >= 0 && a > A.eps_pivot) {
                Key c;
                c.v = rhs / a;
                c.lab = lab;
                c.pos = (int32_t)i;
                k = key_min<false>(k, c);
            }
        }
    }
    const Key lw = cluster_key_min<false>(cluster, k, &slot_ratio, sk, &bc);
    cluster.sync();  // all remote shared-memory reads and all reads of the old state are done
    if (!leader) return;
    const int r = lw.pos;
    if (r < 0) {
        st->pend = 0;
        st->done = 1;
        st->status = 3;  // UNBOUNDED
        st->have_pivot = 0;
        return;
    }
    const double p = __ldcg(colT + r);
    const double inv_p = 1.0 / p;
    st->s = s;
    st->enter_lab = win.lab;
    st->best_val = win.v;
    st->have_pivot = 1;
    st->r = r;
    st->p = p;
    st->inv_p = inv_p;
    if (BLOCKED) {
        A.B.pend->r[t] = r;
        A.B.pend->s[t] = s;
        A.B.pend->inv_p[t] = inv_p;
        A.B.pend->q_rhs = __ldcg(A.B.rhscur + r) / p;
    }
    st->pend = 1;  // rank-1: row r awaits its scaling; look-ahead: the row part of this pivot is due
    const int32_t leave = A.rowlab[r];
    st->leave_lab = leave;
    A.rowlab[r] = win.lab;
    A.collab[s] = leave;
    if (n_piv < A.hist_cap) {
        A.h_row[n_piv] = r;
        A.h_col[n_piv] = s;
        A.h_enter[n_piv] = win.lab;
        A.h_leave[n_piv] = leave;
    }
    st->n_pivots = n_piv + 1;
}

}  // namespace b200lp
