// kernels_blocked.cuh -- look-ahead pivoting: K pivots are DECIDED before the tableau is touched, then applied in
// ONE pass (one read + one write of every element per K pivots instead of per pivot).
//
// A rank-1 pivot loop is bound by HBM at 2*R*C*8 bytes per pivot.  The decisions of a pivot, however, need only
// O(R + C) numbers: the objective row, the right-hand side, the entering column and the pivot row.  With t pivots
// pending (decided, not yet applied), the current value of any element (i, j) is obtained from the stored tableau by
// replaying the t steps on that element alone:
//
//     step u:   i == r_u  ->  q_u[j]                              (row r_u became the scaled pivot row)
//               j == s_u  ->  fma(-col_u[i], 1/p_u, 0)            (column s_u became the leaving variable's column)
//               else      ->  fma(-col_u[i], q_u[j], value)
//
// which is, operation for operation, what the sequential algorithm does to that element.  So every decision and every
// stored number is BIT-IDENTICAL to the rank-1 loop (and to the oracle); only the order in which elements are visited
// changes.  Per pending pivot the state is col_u (R doubles, the entering column as it was) and q_u (C doubles, the
// scaled pivot row); the objective row and the right-hand side are kept current in two small vectors.  Look-ahead pivot
// t+1 costs O(t * (R + C)) flops in three small kernels; the flush applies K steps per element in registers.
#pragma once
#include "common.cuh"

namespace b200lp {

constexpr int BLK_KMAX = 32;
constexpr int BLK_THREADS = 256;

// The number of pending (decided, unapplied) pivots is implicit: DevState.n_pivots - base.
struct BlkPending {
    long long base;          // DevState.n_pivots at the last flush
    double q_rhs;            // rhs_cur[r] / p of the pivot being recorded (set by k_blk_ratio, used by k_blk_row)
    int32_t r[BLK_KMAX];
    int32_t s[BLK_KMAX];
    double inv_p[BLK_KMAX];
};

struct BlkBuffers {
    BlkPending* pend;
    double* colP;    // [KMAX][Rpad]  entering columns as they were when chosen
    double* qP;      // [KMAX][Cpad]  scaled pivot rows
    double* objcur;  // [C] current objective row
    double* rhscur;  // [R] current right-hand side
    int64_t Rpad, Cpad;
};

__device__ __forceinline__ double blk_step(double val, bool is_r, bool is_s, double c_i, double q_j, double inv_p) {
    if (is_r) return q_j;
    return __fma_rn(-c_i, is_s ? inv_p : q_j, is_s ? 0.0 : val);
}

// Replay of the t pending steps on ONE element (i, j).  One operand of every step comes from global memory (g[u *
// gstride]: q_u[j] when G_IS_Q, else col_u[i]), the other from shared memory (sm[u]).  BATCH: the global operands are
// loaded 8 at a time ahead of the chain so that their L2 latencies overlap -- right for k_blk_flush_special (many CTAs,
// one element per thread, latency-bound); measured slower in the picks, which are bound by the L2 bandwidth of the 16
// SMs of the cluster (t * (R + C) * 8 bytes per pick), not by the latency of the individual loads.
template <bool G_IS_Q, bool BATCH = false>
__device__ __forceinline__ double blk_replay(double v, int t, const double* __restrict__ g, int64_t gstride,
                                             const double* sm, const int32_t* sr, const int32_t* ss,
                                             const double* sinv, int64_t i, int64_t j) {
    if (BATCH) {
        for (int u0 = 0; u0 < t; u0 += 8) {
            double gv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[k] = (u0 + k < t) ? g[(int64_t)(u0 + k) * gstride] : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int u = u0 + k;
                if (u < t)
                    v = blk_step(v, i == sr[u], j == ss[u], G_IS_Q ? sm[u] : gv[k], G_IS_Q ? gv[k] : sm[u], sinv[u]);
            }
        }
    } else {
        for (int u = 0; u < t; ++u) {
            const double gv = g[(int64_t)u * gstride];
            v = blk_step(v, i == sr[u], j == ss[u], G_IS_Q ? sm[u] : gv, G_IS_Q ? gv : sm[u], sinv[u]);
        }
    }
    return v;
}

// copy the objective row and the right-hand side out of the stored tableau; no pivots pending
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_init(const double* __restrict__ T, int64_t R, int64_t C, int64_t ld, int64_t obj_row, const DevState* st,
           BlkBuffers B) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = tid; j < C; j += n) B.objcur[j] = T[obj_row * ld + j];
    for (int64_t i = tid; i < R; i += n) B.rhscur[i] = T[i * ld + C - 1];
    if (tid == 0) B.pend->base = st->n_pivots;
}

// Ratio test of the look-ahead pivot: the entering column is gathered from the stored tableau and brought up to date
// by replaying the pending steps; it is kept as col_t.  Same ticket pattern and bookkeeping as k_ratio.
// `ext` (sharded tableaux): the gathered candidates; the winner's column arrives already up to date, so it is copied
// instead of gathered + replayed, and st->s is -1 on the shards that do not own it.
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_ratio(const double* __restrict__ T, int64_t R, int64_t m, int64_t C, int64_t ld, int32_t* rowlab, int32_t* collab,
            double eps_pivot, DevState* st, Key* partials, BlkBuffers B, const double* __restrict__ ext,
            int64_t ext_stride, int32_t* h_row, int32_t* h_col, int32_t* h_enter, int32_t* h_leave, int64_t hist_cap) {
    __shared__ Key sk[BLK_THREADS / 32];
    __shared__ bool is_last;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sq[BLK_KMAX];
    if (st->done || !st->have_pivot) return;
    const int s = st->s;
    const int t = (int)(st->n_pivots - B.pend->base);  // this kernel's last CTA increments n_pivots at its very end
    const double* src = ext ? ext + (int64_t)st->win_rank * ext_stride + 2 : nullptr;
    if (!src && threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
        sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
        sq[threadIdx.x] = B.qP[(int64_t)threadIdx.x * B.Cpad + s];  // q_u[s]
    }
    __syncthreads();
    double* colT = B.colP + (int64_t)t * B.Rpad;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    Key k = key_none();
    for (int64_t i = tid; i < R; i += nthr) {
        double a;
        if (src) {
            a = src[i];
        } else {
            a = T[i * ld + s];
            a = blk_replay<false>(a, t, B.colP + i, B.Rpad, sq, sr, ss, sinv, i, s);
        }
        colT[i] = a;
        if (i < m) {
            const int32_t lab = rowlab[i];
            if (lab >= 0 && a > eps_pivot) {
                Key c;
                c.v = B.rhscur[i] / a;
                c.lab = lab;
                c.pos = (int32_t)i;
                k = key_min<false>(k, c);
            }
        }
    }
    k = block_key_min<false>(k, sk);
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = k;
            __threadfence();
            const unsigned int tk = atomicAdd(&st->ticket_ratio, 1u);
            is_last = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        k = key_none();
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            Key c;
            c.v = __ldcg(&partials[b].v);
            c.lab = __ldcg(&partials[b].lab);
            c.pos = __ldcg(&partials[b].pos);
            k = key_min<false>(k, c);
        }
        __syncthreads();
        k = block_key_min<false>(k, sk);
    }
    if (threadIdx.x == 0) {
        st->ticket_ratio = 0;
        const int r = k.pos;
        if (r < 0) {
            st->done = 1;
            st->status = 3;  // UNBOUNDED
            st->have_pivot = 0;
        } else {
            const double p = __ldcg(colT + r);
            const double inv_p = 1.0 / p;
            st->r = r;
            st->p = p;
            st->inv_p = inv_p;
            B.pend->r[t] = r;
            B.pend->s[t] = s;
            B.pend->inv_p[t] = inv_p;
            B.pend->q_rhs = B.rhscur[r] / p;
            const int32_t leave = rowlab[r];
            st->leave_lab = leave;
            rowlab[r] = st->enter_lab;
            if (s >= 0) collab[s] = leave;
            const long long n = st->n_pivots;
            if (n < hist_cap) {
                h_row[n] = r;
                h_col[n] = s;
                h_enter[n] = st->enter_lab;
                h_leave[n] = leave;
            }
            st->n_pivots = n + 1;
            st->pend = 1;  // look-ahead loops: "the row part of this pivot is due" (cleared by whoever does it)
        }
    }
}

// Pivot row of the look-ahead pivot brought up to date and scaled (q_t), then the objective row and the right-hand
// side advanced by this pivot (k_blk_ratio already counted it: its index among the pending ones is n_pivots-1-base).
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_row(const double* __restrict__ T, int64_t R, int64_t C, int64_t ld, int64_t obj_row, DevState* st, BlkBuffers B) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sc[BLK_KMAX];
    if (!st->pend) return;  // the flag is cleared by k_blk_clear, after the flush
    const int t = (int)(st->n_pivots - 1 - B.pend->base);
    const int r = st->r, s = st->s;
    const double p = st->p, inv_p = st->inv_p;
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
        sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
        sc[threadIdx.x] = B.colP[(int64_t)threadIdx.x * B.Rpad + r];  // col_u[r]
    }
    __syncthreads();
    const double* colT = B.colP + (int64_t)t * B.Rpad;
    double* qT = B.qP + (int64_t)t * B.Cpad;
    const double c_obj = colT[obj_row];
    const double q_rhs = B.pend->q_rhs;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = tid; j < C; j += nthr) {
        double v = T[(int64_t)r * ld + j];
        v = blk_replay<true>(v, t, B.qP + j, B.Cpad, sc, sr, ss, sinv, r, j);
        const double q = (j == s) ? inv_p : v / p;
        qT[j] = q;
        B.objcur[j] = blk_step(B.objcur[j], false, j == s, c_obj, q, inv_p);
    }
    for (int64_t i = tid; i < R; i += nthr)
        B.rhscur[i] = (i == r) ? q_rhs : __fma_rn(-colT[i], q_rhs, B.rhscur[i]);
}

// Fused "row part of the previous look-ahead pivot" + "pricing of the next one": the thread that refreshes objcur[j]
// prices it at once, so a look-ahead pivot costs two launches (this one and k_blk_ratio) instead of three.
// DevState.pend == 1 on entry means k_blk_ratio recorded a pivot whose row part is still due.  SHARDED: "no local
// candidate" does not end the loop (k_shard_winner decides after the all-gather).
template <bool BLAND, bool SHARDED>
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_rowprice(const double* __restrict__ T, int64_t R, int64_t C, int64_t ld, int64_t obj_row,
               const int32_t* __restrict__ collab, int32_t art_base, double eps_cost, DevState* st, Key* partials,
               BlkBuffers B) {
    __shared__ Key sk[BLK_THREADS / 32];
    __shared__ bool is_last;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sc[BLK_KMAX];
    if (st->done) return;
    const bool row_due = st->pend != 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    Key k = key_none();
    if (row_due) {
        const int t = (int)(st->n_pivots - 1 - B.pend->base);
        const int r = st->r, s = st->s;
        const double p = st->p, inv_p = st->inv_p;
        if (threadIdx.x < t) {
            sr[threadIdx.x] = B.pend->r[threadIdx.x];
            ss[threadIdx.x] = B.pend->s[threadIdx.x];
            sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
            sc[threadIdx.x] = B.colP[(int64_t)threadIdx.x * B.Rpad + r];
        }
        __syncthreads();
        const double* colT = B.colP + (int64_t)t * B.Rpad;
        double* qT = B.qP + (int64_t)t * B.Cpad;
        const double c_obj = colT[obj_row];
        const double q_rhs = B.pend->q_rhs;
        for (int64_t j = tid; j < C; j += nthr) {
            double v = T[(int64_t)r * ld + j];
            v = blk_replay<true>(v, t, B.qP + j, B.Cpad, sc, sr, ss, sinv, r, j);
            const double q = (j == s) ? inv_p : v / p;
            qT[j] = q;
            const double d = blk_step(B.objcur[j], false, j == s, c_obj, q, inv_p);
            B.objcur[j] = d;
            if (j < C - 1) {
                // labels were swapped by k_blk_ratio already: collab[s] is the variable that just left the basis
                const int32_t lab = collab[j];
                if (lab < art_base && d < -eps_cost) {
                    Key c;
                    c.v = d;
                    c.lab = lab;
                    c.pos = (int32_t)j;
                    k = key_min<BLAND>(k, c);
                }
            }
        }
        for (int64_t i = tid; i < R; i += nthr)
            B.rhscur[i] = (i == r) ? q_rhs : __fma_rn(-colT[i], q_rhs, B.rhscur[i]);
    } else {
        for (int64_t j = tid; j < C - 1; j += nthr) {
            const int32_t lab = collab[j];
            const double d = B.objcur[j];
            if (lab < art_base && d < -eps_cost) {
                Key c;
                c.v = d;
                c.lab = lab;
                c.pos = (int32_t)j;
                k = key_min<BLAND>(k, c);
            }
        }
    }
    k = block_key_min<BLAND>(k, sk);
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = k;
            __threadfence();
            const unsigned int tk = atomicAdd(&st->ticket_price, 1u);
            is_last = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        k = key_none();
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            Key c;
            c.v = __ldcg(&partials[b].v);
            c.lab = __ldcg(&partials[b].lab);
            c.pos = __ldcg(&partials[b].pos);
            k = key_min<BLAND>(k, c);
        }
        __syncthreads();
        k = block_key_min<BLAND>(k, sk);
    }
    if (threadIdx.x == 0) {
        st->ticket_price = 0;
        st->pend = 0;  // the row part (if any) is done
        if (st->n_pivots >= st->max_pivots) {
            st->done = 1;
            st->status = 1;  // LIMIT
            st->have_pivot = 0;
        } else if (k.lab == B200LP_NO_LAB) {
            st->have_pivot = 0;
            st->s = -1;
            st->enter_lab = -1;
            if (!SHARDED) {  // sharded: another shard may still have a candidate (k_shard_winner decides)
                st->done = 1;
                st->status = 0;  // OPTIMAL for this objective row
            }
        } else {
            st->have_pivot = 1;
            st->s = k.pos;
            st->enter_lab = k.lab;
            st->best_val = k.v;
        }
    }
}

// ---- the flush: k_blk_flush_special + k_blk_flush_db ----------------------------------------------------------------
// Every element of the stored tableau replays the t pending steps.  Only the elements in a pivot row or a pivot column of
// a pending step need the general replay (blk_step); there are at most K rows and K columns of them.
// k_blk_flush_special replays exactly those (pivot rows; pivot columns together with the other column of their 16-byte
// pair), in place.  k_blk_flush_db replays everything else with the plain chain
//     v <- fma(-col_u[i], q_u[j], v),  u = 0 .. t-1
// and neither loads nor stores the special rows / column pairs, so it has no divergent slow path and no tile is slower
// than the others (with the slow path inside the main kernel, the 12 % of the tiles that hold a pivot row set its
// duration).  The two kernels touch disjoint elements and an element's replay reads only its own old value, col_u[i] and
// q_u[j], so their order is free.
// Thread = one column (row part) or one row (column part): its t history values -- q_u[j] or col_u[i] -- are loaded
// once into registers and reused for every distinct pivot row / pivot column pair, so the kernel reads the history once
// (t * (R + C) * 8 bytes) instead of once per pivot row and column.  blockIdx.y: 0 = rows, 1 = column pairs.
__global__ void __launch_bounds__(256)
k_blk_flush_special(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* st, BlkBuffers B) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX];
    __shared__ int32_t list[BLK_KMAX];               // distinct pivot rows / first columns of the distinct pairs
    __shared__ int n_list;
    __shared__ double sm[BLK_KMAX][BLK_KMAX];        // rows: col_u[r_k];  columns: q_u[j0_k]
    __shared__ double sm2[BLK_KMAX][BLK_KMAX];       // columns: q_u[j0_k + 1]
    const int t = (int)(st->n_pivots - B.pend->base);
    if (t == 0) return;
    const bool cols = blockIdx.y != 0;
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
        sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int u = 0; u < t; ++u) {
            // (sharded tableaux: ss[u] < 0 when the entering column of step u lives in another shard)
            const int key = cols ? (ss[u] < 0 ? -1 : (ss[u] & ~1)) : sr[u];
            bool seen = key < 0;
            for (int k = 0; k < n && !seen; ++k) seen = list[k] == key;
            if (!seen) list[n++] = key;
        }
        n_list = n;
    }
    __syncthreads();
    const int nl = n_list;
    if (nl == 0) return;
    for (int e = threadIdx.x; e < nl * t; e += blockDim.x) {
        const int k = e / t, u = e - k * t;
        if (!cols) {
            sm[k][u] = B.colP[(int64_t)u * B.Rpad + list[k]];
        } else {
            sm[k][u] = B.qP[(int64_t)u * B.Cpad + list[k]];
            sm2[k][u] = B.qP[(int64_t)u * B.Cpad + list[k] + 1];  // (qP is padded past C)
        }
    }
    __syncthreads();
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double h[BLK_KMAX];  // this thread's history: q_u[j] (rows) or col_u[i] (columns)
    if (!cols) {
        const int64_t j = x;
        if (j >= C) return;
#pragma unroll
        for (int u = 0; u < BLK_KMAX; ++u) h[u] = u < t ? B.qP[(int64_t)u * B.Cpad + j] : 0.0;
        for (int k = 0; k < nl; ++k) {
            const int r = list[k];
            double v = T[(int64_t)r * ld + j];
#pragma unroll
            for (int u = 0; u < BLK_KMAX; ++u)
                if (u < t) v = blk_step(v, r == sr[u], j == ss[u], sm[k][u], h[u], sinv[u]);
            T[(int64_t)r * ld + j] = v;
        }
    } else {
        const int64_t i = x;
        if (i >= R) return;
        bool pivot_row = false;
        for (int u = 0; u < t; ++u) pivot_row |= (i == sr[u]);
        if (pivot_row) return;  // done by the row part
#pragma unroll
        for (int u = 0; u < BLK_KMAX; ++u) h[u] = u < t ? B.colP[(int64_t)u * B.Rpad + i] : 0.0;
        for (int k = 0; k < nl; ++k) {
            const int64_t j0 = list[k];
            const bool two = j0 + 1 < C;
            double vx = T[i * ld + j0], vy = two ? T[i * ld + j0 + 1] : 0.0;
#pragma unroll
            for (int u = 0; u < BLK_KMAX; ++u)
                if (u < t) {
                    vx = blk_step(vx, false, j0 == ss[u], h[u], sm[k][u], sinv[u]);
                    vy = blk_step(vy, false, j0 + 1 == ss[u], h[u], sm2[k][u], sinv[u]);
                }
            T[i * ld + j0] = vx;
            if (two) T[i * ld + j0 + 1] = vy;
        }
    }
}

// Main kernel.  WC warps across by 8 / WC down; tile = TR rows x one strip of SW columns; a CTA takes a contiguous range
// of tiles in strip-major order, i.e. walks DOWN a strip, so q_u of the strip (SW columns x K steps) is staged in shared
// memory once per strip and read per step; a warp's row group is 8 rows x its columns, all t steps in registers.
// Measured on 16384^2 at K = 32 (scripts/probe_lookahead.py): 2.14 ms for the first version of the flush (tiles of 64 x 512
// taken grid-stride, q_u re-read per row group through L1 with a 25 % hit rate, slow path per tile) -> 1.10 ms (special
// kernel + strips + q in shared memory) -> 0.86 ms with the two software pipelines below; what is left is the FP64 pipe
// (16 us per step = 93 % of its peak) plus the part of the HBM time the replay does not cover.
//  * tableau rows: a warp issues the loads of its NEXT row group right after step 0 of the current one (registers v /
//    vn), so the HBM latency is covered by its own replay instead of stalling all warps of the SM in a convoy;
//  * col_u slices: the next tile's slices are copied global -> shared with cp.async into the other half of a double
//    buffer while the current tile is replayed; one __syncthreads per tile.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// NP = 16-byte column pairs per thread (pair p of a lane lies 64 * p columns to the right, so every 128-bit access of a
// warp is one contiguous 512-byte segment).  NP = 1: 2 CTAs per SM.  NP = 2: a col_u value read from shared memory feeds
// 4 FMAs instead of 2 -- the shared-memory pipe is the co-bottleneck (72 % of its peak at 45 % FP64 with NP = 1) -- at
// the price of 128 registers for v and vn alone, hence 1 CTA per SM.
template <int WC, int TR, int NP>
__global__ void __launch_bounds__(256, NP == 1 ? 2 : 1)
k_blk_flush_db(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* st, BlkBuffers B, int64_t n_rb,
               int64_t n_strips) {
    constexpr int WR = 8 / WC;
    constexpr int WW = 64 * NP;   // columns per warp
    constexpr int SW = WW * WC;   // columns per strip
    constexpr int NG = TR / (8 * WR);
    static_assert(TR % (8 * WR) == 0 && TR % 2 == 0, "tile rows");
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    extern __shared__ __align__(16) double dyn[];
    double* scol = dyn;                         // [2][BLK_KMAX][TR]
    double* sq = dyn + 2 * BLK_KMAX * TR;       // [t][SW]
    const int t = (int)(st->n_pivots - B.pend->base);
    if (t == 0) return;
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wc = warp % WC, wr = warp / WC;
    const int64_t n_items = n_rb * n_strips;
    const int64_t it0 = n_items * blockIdx.x / gridDim.x, it1 = n_items * (blockIdx.x + 1) / gridDim.x;
    if (it0 >= it1) return;

    // asynchronous copy of the col_u slices of row block rb into half `buf` (16-byte pieces; Rpad and TR are even)
    auto stage_cols = [&](int64_t rb, int buf) {
        const int64_t i0 = rb * TR;
        const int rows = (int)(min(R, i0 + (int64_t)TR) - i0);
        const int pieces = (rows + 1) / 2;  // colP is padded to Rpad >= R + (R & 1)
        double* dst = scol + (int64_t)buf * BLK_KMAX * TR;
        for (int e = threadIdx.x; e < t * (TR / 2); e += 256) {
            const int u = e / (TR / 2), pc = e - u * (TR / 2);
            if (pc < pieces) cp_async16(dst + u * TR + 2 * pc, B.colP + (int64_t)u * B.Rpad + i0 + 2 * pc);
        }
        cp_async_commit();
    };

    // this warp's walk over its row groups: (strip, row block, group) of the group whose loads are issued next
    int64_t n_strip = it0 / n_rb, n_rbk = it0 - n_strip * n_rb, n_it = it0;
    int n_g = 0;
    int64_t skip_strip = -1;
    bool skip[NP];
    struct Grp {
        double* base;
        unsigned off[NP];  // bit w: do not touch row w of pair p (pivot row / column pair, beyond the tableau)
    };
    auto next_group = [&]() {
        Grp G;
        const int64_t row0 = n_rbk * TR + (n_g * WR + wr) * 8;
        const int64_t jj = n_strip * SW + wc * WW + 2 * lane;
        unsigned off = 0;
        for (int u = lane; u < t; u += 32) {
            const int64_t d = (int64_t)sr[u] - row0;
            if (d >= 0 && d < 8) off |= 1u << (int)d;
        }
        off = __reduce_or_sync(0xffffffffu, off);
        if (R - row0 < 8) off |= (R - row0 <= 0) ? 0xffu : (0xffu << (int)(R - row0)) & 0xffu;
        if (n_strip != skip_strip) {  // a CTA changes strip at most a few times
            skip_strip = n_strip;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                skip[p] = jj + 64 * p >= C;
                for (int u = 0; u < t; ++u) skip[p] |= (ss[u] >= 0 && (ss[u] & ~1) == jj + 64 * p);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) G.off[p] = (n_it >= it1 || skip[p]) ? 0xffu : off;
        G.base = T + row0 * ld + jj;
        if (++n_g == NG) {
            n_g = 0;
            ++n_it;
            if (++n_rbk == n_rb) {
                n_rbk = 0;
                ++n_strip;
            }
        }
        return G;
    };
    auto load = [&](const Grp& G, double2 (*v)[NP]) {
#pragma unroll
        for (int w = 0; w < 8; ++w)
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                v[w][p] = make_double2(0.0, 0.0);
                if (!((G.off[p] >> w) & 1u))
                    v[w][p] = ld_stream(reinterpret_cast<const double2*>(G.base + (int64_t)w * ld + 64 * p));
            }
    };

    double2 vn[8][NP];
    Grp gn = next_group();
    load(gn, vn);
    int64_t strip = it0 / n_rb, rb = it0 - strip * n_rb, cur_strip = -1;
    stage_cols(rb, 0);
    for (int64_t it = it0; it < it1; ++it) {
        const int buf = (int)((it - it0) & 1);
        cp_async_wait_all();
        __syncthreads();  // this tile's col_u slices have landed; everybody is done with the previous tile
        if (strip != cur_strip) {
            cur_strip = strip;
            for (int u = warp; u < t; u += 8)
                for (int c = lane; c < SW; c += 32) {
                    const int64_t jj = strip * SW + c;
                    sq[u * SW + c] = jj < B.Cpad ? B.qP[(int64_t)u * B.Cpad + jj] : 0.0;
                }
            __syncthreads();
        }
        // next tile of this CTA (same strip, next row block, or the top of the next strip)
        int64_t nstrip = strip, nrb = rb + 1;
        if (nrb == n_rb) {
            nrb = 0;
            ++nstrip;
        }
        if (it + 1 < it1) stage_cols(nrb, buf ^ 1);
        const double* qa = sq + wc * WW + 2 * lane;
        const double* sc_tile = scol + (int64_t)buf * BLK_KMAX * TR;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            const int rr = (g * WR + wr) * 8;  // rows past the tile's end are masked in `off` and computed on zeros
            const Grp gc = gn;
            double2 v[8][NP];
#pragma unroll
            for (int w = 0; w < 8; ++w)
#pragma unroll
                for (int p = 0; p < NP; ++p) v[w][p] = vn[w][p];
            auto step = [&](int u) {
                const double2* sc = reinterpret_cast<const double2*>(sc_tile + u * TR + rr);
                const double2 c01 = sc[0], c23 = sc[1], c45 = sc[2], c67 = sc[3];
                const double c[8] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y, c67.x, c67.y};
                double2 q[NP];
#pragma unroll
                for (int p = 0; p < NP; ++p) q[p] = *reinterpret_cast<const double2*>(qa + u * SW + 64 * p);
#pragma unroll
                for (int w = 0; w < 8; ++w)
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        v[w][p].x = __fma_rn(-c[w], q[p].x, v[w][p].x);
                        v[w][p].y = __fma_rn(-c[w], q[p].y, v[w][p].y);
                    }
            };
            // Step 0 first: it consumes every register of the previous batch of loads, i.e. the wait on their
            // scoreboard happens HERE, before the same load instructions are issued again for the next group (the
            // loads of consecutive iterations share one scoreboard, so a wait placed after the re-issue would wait for
            // the new batch as well and the prefetch would buy nothing).
            step(0);
            asm volatile("" ::: "memory");
            gn = next_group();
            load(gn, vn);
            asm volatile("" ::: "memory");
#pragma unroll 4
            for (int u = 1; u < t; ++u) step(u);
#pragma unroll
            for (int w = 0; w < 8; ++w)
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    if (!((gc.off[p] >> w) & 1u))
                        st_stream(reinterpret_cast<double2*>(gc.base + (int64_t)w * ld + 64 * p), v[w][p]);
        }
        strip = nstrip;
        rb = nrb;
    }
}

// The instance the library launches: 512-column strips, 4 columns per thread, 1 CTA per SM (22.8k pivots/s on 16384^2 at
// K = 32 against 21.5k for <4, 64, 1> with 2 CTAs per SM).
constexpr int FL_WC = 4, FL_TR = 128, FL_NP = 2;
constexpr size_t FL_SMEM_BYTES = (size_t)BLK_KMAX * (64 * FL_NP * FL_WC + 2 * FL_TR) * 8;  // q_u strip + 2 x col_u slices
static_assert(FL_SMEM_BYTES <= 227 * 1024, "flush shared memory");

// after the flush: nothing pending, and the row part of the block's last pivot has been done by k_blk_row
__global__ void k_blk_clear(DevState* st, BlkBuffers B) {
    B.pend->base = st->n_pivots;
    st->pend = 0;
}

// ---- peer-memory exchange of the sharded loops (replaces the caller's all-gather; kernel in kernels_shard.cuh) ------
// Every shard owns an exchange region  [parity 2][world][xstride] doubles + [parity 2][world] 64-bit generation words
// that all its peers can address (NVLink / NVSwitch peer memory; the host passes the peer addresses, e.g. torch
// symmetric memory).  A record is [reduced cost, variable id, column (R doubles)], xstride = R + 2 rounded up to even.
struct P2PPeers {
    double* base[16];   // base[g] = rank g's region as addressed from this GPU
    int32_t world, rank;
    int64_t xstride;    // R + 2, rounded up to even
};

__device__ __forceinline__ unsigned long long* p2p_flags(double* region, int world, int64_t xstride) {
    return reinterpret_cast<unsigned long long*>(region + 2 * (int64_t)world * xstride);
}

// Sharded look-ahead: publish this shard's candidate [reduced cost, variable id, column brought up to date].
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_shard_extract(const double* __restrict__ T, int64_t R, int64_t ld, const DevState* st, BlkBuffers B,
                    double* __restrict__ cand) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sq[BLK_KMAX];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int have = st->have_pivot && !st->done;
    if (tid == 0) {
        cand[0] = have ? st->best_val : 0.0;
        cand[1] = have ? (double)st->enter_lab : -1.0;
    }
    if (!have) return;
    const int s = st->s;
    const int t = (int)(st->n_pivots - B.pend->base);
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
        sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
        sq[threadIdx.x] = B.qP[(int64_t)threadIdx.x * B.Cpad + s];
    }
    __syncthreads();
    for (int64_t i = tid; i < R; i += (int64_t)gridDim.x * blockDim.x) {
        double a = T[i * ld + s];
        a = blk_replay<false>(a, t, B.colP + i, B.Rpad, sq, sr, ss, sinv, i, s);
        cand[2 + i] = a;
    }
}

}  // namespace b200lp
