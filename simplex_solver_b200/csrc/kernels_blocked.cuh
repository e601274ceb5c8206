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

constexpr int BLK_KMAX = 16;
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
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_ratio(const double* __restrict__ T, int64_t R, int64_t m, int64_t C, int64_t ld, int32_t* rowlab, int32_t* collab,
            double eps_pivot, DevState* st, Key* partials, BlkBuffers B, int32_t* h_row, int32_t* h_col,
            int32_t* h_enter, int32_t* h_leave, int64_t hist_cap) {
    __shared__ Key sk[BLK_THREADS / 32];
    __shared__ bool is_last;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sq[BLK_KMAX];
    if (st->done || !st->have_pivot) return;
    const int s = st->s;
    const int t = (int)(st->n_pivots - B.pend->base);  // this kernel's last CTA increments n_pivots at its very end
    if (threadIdx.x < t) {
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
        double a = T[i * ld + s];
        for (int u = 0; u < t; ++u)
            a = blk_step(a, i == sr[u], s == ss[u], B.colP[(int64_t)u * B.Rpad + i], sq[u], sinv[u]);
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
            collab[s] = leave;
            const long long n = st->n_pivots;
            if (n < hist_cap) {
                h_row[n] = r;
                h_col[n] = s;
                h_enter[n] = st->enter_lab;
                h_leave[n] = leave;
            }
            st->n_pivots = n + 1;
        }
    }
}

// Pivot row of the look-ahead pivot brought up to date and scaled (q_t), then the objective row and the right-hand
// side advanced by this pivot (k_blk_ratio already counted it: its index among the pending ones is n_pivots-1-base).
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_row(const double* __restrict__ T, int64_t R, int64_t C, int64_t ld, int64_t obj_row, DevState* st, BlkBuffers B) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sc[BLK_KMAX];
    if (st->done || !st->have_pivot) return;
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
        for (int u = 0; u < t; ++u)
            v = blk_step(v, r == sr[u], j == ss[u], sc[u], B.qP[(int64_t)u * B.Cpad + j], sinv[u]);
        const double q = (j == s) ? inv_p : v / p;
        qT[j] = q;
        B.objcur[j] = blk_step(B.objcur[j], false, j == s, c_obj, q, inv_p);
    }
    for (int64_t i = tid; i < R; i += nthr)
        B.rhscur[i] = (i == r) ? q_rhs : __fma_rn(-colT[i], q_rhs, B.rhscur[i]);
}

// The flush: every element of the stored tableau replays the pending steps in registers.  Same tiling as
// k_update_ldg (256 threads x 2 columns x up to 64 rows, streaming 128-bit loads/stores, 8 rows in flight per thread);
// q_u for the thread's two columns lives in registers for all pending u; the tile's slice of every col_u is staged in
// shared memory once per tile and read as a broadcast.
constexpr int BLK_TILE_ROWS = 64;
constexpr int BLK_UNROLL = 8;

template <int KMAX>
__global__ void __launch_bounds__(256, (KMAX <= 8 ? 2 : 1))
k_blk_flush(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* st, BlkBuffers B, int tile_rows,
            int tiles_c, int64_t n_tiles) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX];
    __shared__ double scol[KMAX][BLK_TILE_ROWS];
    const int t = (int)(st->n_pivots - B.pend->base);
    if (t == 0) return;
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
        sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tc = tile % tiles_c, tr = tile / tiles_c;
        const int64_t i0 = tr * tile_rows;
        const int rows = (int)(min(R, i0 + (int64_t)tile_rows) - i0);
        __syncthreads();  // previous tile's readers are done with scol (and sr/ss/sinv are written)
        for (int e = threadIdx.x; e < t * rows; e += blockDim.x) {
            const int u = e / rows, rr = e - u * rows;
            scol[u][rr] = B.colP[(int64_t)u * B.Rpad + i0 + rr];
        }
        __syncthreads();
        const int64_t j = tc * 512 + 2 * threadIdx.x;
        if (j >= C) continue;
        double qx[KMAX], qy[KMAX];
        unsigned mx = 0, my = 0;  // bit u: this thread's column is s_u
#pragma unroll
        for (int u = 0; u < KMAX; ++u) {
            qx[u] = qy[u] = 0.0;
            if (u < t) {
                const double2 q = *reinterpret_cast<const double2*>(B.qP + (int64_t)u * B.Cpad + j);
                qx[u] = q.x;
                qy[u] = q.y;
                if (j == ss[u]) mx |= 1u << u;
                if (j + 1 == ss[u]) my |= 1u << u;
            }
        }
        // does this tile hold a pivot row, or this thread a pivot column?  If not, the replay is t plain FMAs.
        bool special_rows = false;
        for (int u = 0; u < t; ++u) special_rows |= (sr[u] >= i0 && sr[u] < i0 + rows);
        const bool plain = !special_rows && mx == 0 && my == 0;
        double* base = T + i0 * ld + j;
        for (int rr = 0; rr < rows; rr += BLK_UNROLL) {
            double2 v[BLK_UNROLL];
#pragma unroll
            for (int w = 0; w < BLK_UNROLL; ++w)
                if (rr + w < rows) v[w] = ld_stream(reinterpret_cast<const double2*>(base + (int64_t)(rr + w) * ld));
            if (plain) {
#pragma unroll
                for (int u = 0; u < KMAX; ++u) {
                    if (u < t) {
#pragma unroll
                        for (int w = 0; w < BLK_UNROLL; ++w) {
                            const double nc = -scol[u][(rr + w) & (BLK_TILE_ROWS - 1)];
                            v[w].x = __fma_rn(nc, qx[u], v[w].x);
                            v[w].y = __fma_rn(nc, qy[u], v[w].y);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < KMAX; ++u) {
                    if (u < t) {
#pragma unroll
                        for (int w = 0; w < BLK_UNROLL; ++w) {
                            const double c = scol[u][(rr + w) & (BLK_TILE_ROWS - 1)];
                            const bool is_r = (i0 + rr + w == sr[u]);
                            v[w].x = blk_step(v[w].x, is_r, (mx >> u) & 1u, c, qx[u], sinv[u]);
                            v[w].y = blk_step(v[w].y, is_r, (my >> u) & 1u, c, qy[u], sinv[u]);
                        }
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < BLK_UNROLL; ++w)
                if (rr + w < rows) st_stream(reinterpret_cast<double2*>(base + (int64_t)(rr + w) * ld), v[w]);
        }
    }
}
__global__ void k_blk_clear(const DevState* st, BlkBuffers B) { B.pend->base = st->n_pivots; }

}  // namespace b200lp
