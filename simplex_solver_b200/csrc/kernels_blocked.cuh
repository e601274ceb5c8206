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
            for (int u = 0; u < t; ++u)
                a = blk_step(a, i == sr[u], s == ss[u], B.colP[(int64_t)u * B.Rpad + i], sq[u], sinv[u]);
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
        for (int u = 0; u < t; ++u)
            v = blk_step(v, r == sr[u], j == ss[u], sc[u], B.qP[(int64_t)u * B.Cpad + j], sinv[u]);
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
            for (int u = 0; u < t; ++u)
                v = blk_step(v, r == sr[u], j == ss[u], sc[u], B.qP[(int64_t)u * B.Cpad + j], sinv[u]);
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

// The flush: every element of the stored tableau replays the pending steps in registers.  Same tiling as
// k_update_ldg (256 threads x 2 columns x up to 64 rows, streaming 128-bit loads/stores, 8 rows in flight per thread).
// The steps are taken in chunks of 8: q_u of the chunk for the thread's two columns is (re)loaded into registers per
// row group (an L1/L2 hit), so the register footprint -- and with it the occupancy that hides the HBM latency -- does
// not grow with K.  The tile's slice of every col_u is staged in shared memory once per tile (broadcast reads).
constexpr int BLK_TILE_ROWS = 64;
constexpr int BLK_UNROLL = 8;
constexpr int BLK_CHUNK = 8;

__global__ void __launch_bounds__(256, 2)
k_blk_flush(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* st, BlkBuffers B, int tile_rows,
            int tiles_c, int64_t n_tiles) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX];
    __shared__ __align__(16) double scol[BLK_KMAX][BLK_TILE_ROWS];
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
        unsigned mx = 0, my = 0;  // bit u: this thread's column is s_u
        bool special_rows = false;
        for (int u = 0; u < t; ++u) {
            if (j == ss[u]) mx |= 1u << u;
            if (j + 1 == ss[u]) my |= 1u << u;
            special_rows |= (sr[u] >= i0 && sr[u] < i0 + rows);
        }
        // does this tile hold a pivot row, or this thread a pivot column?  If not, the replay is t plain FMAs.
        const bool plain = !special_rows && mx == 0 && my == 0;
        const double* qbase = B.qP + j;
        double* base = T + i0 * ld + j;
        for (int rr = 0; rr < rows; rr += BLK_UNROLL) {
            double2 v[BLK_UNROLL];
#pragma unroll
            for (int w = 0; w < BLK_UNROLL; ++w)
                if (rr + w < rows) v[w] = ld_stream(reinterpret_cast<const double2*>(base + (int64_t)(rr + w) * ld));
            for (int u0 = 0; u0 < t; u0 += BLK_CHUNK) {
                double2 q[BLK_CHUNK];
#pragma unroll
                for (int k = 0; k < BLK_CHUNK; ++k)
                    if (u0 + k < t) q[k] = *reinterpret_cast<const double2*>(qbase + (int64_t)(u0 + k) * B.Cpad);
                if (plain && u0 + BLK_CHUNK <= t) {
                    // full chunk, no pivot row / column in sight: 4 x LDS.128 + 16 x DFMA per step for the 8 rows
                    static_assert(BLK_UNROLL == 8, "the vectorised column reads below assume 8 rows per group");
#pragma unroll
                    for (int k = 0; k < BLK_CHUNK; ++k) {
                        const double2* sc = reinterpret_cast<const double2*>(&scol[u0 + k][rr]);
                        const double2 c01 = sc[0], c23 = sc[1], c45 = sc[2], c67 = sc[3];
                        const double c[8] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y, c67.x, c67.y};
#pragma unroll
                        for (int w = 0; w < BLK_UNROLL; ++w) {
                            v[w].x = __fma_rn(-c[w], q[k].x, v[w].x);
                            v[w].y = __fma_rn(-c[w], q[k].y, v[w].y);
                        }
                    }
                } else if (plain) {
#pragma unroll
                    for (int k = 0; k < BLK_CHUNK; ++k) {
                        if (u0 + k < t) {
#pragma unroll
                            for (int w = 0; w < BLK_UNROLL; ++w) {
                                const double nc = -scol[u0 + k][(rr + w) & (BLK_TILE_ROWS - 1)];
                                v[w].x = __fma_rn(nc, q[k].x, v[w].x);
                                v[w].y = __fma_rn(nc, q[k].y, v[w].y);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < BLK_CHUNK; ++k) {
                        if (u0 + k < t) {
                            const int u = u0 + k;
#pragma unroll
                            for (int w = 0; w < BLK_UNROLL; ++w) {
                                const double c = scol[u][(rr + w) & (BLK_TILE_ROWS - 1)];
                                const bool is_r = (i0 + rr + w == sr[u]);
                                v[w].x = blk_step(v[w].x, is_r, (mx >> u) & 1u, c, q[k].x, sinv[u]);
                                v[w].y = blk_step(v[w].y, is_r, (my >> u) & 1u, c, q[k].y, sinv[u]);
                            }
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

// after the flush: nothing pending, and the row part of the block's last pivot has been done by k_blk_row
__global__ void k_blk_clear(DevState* st, BlkBuffers B) {
    B.pend->base = st->n_pivots;
    st->pend = 0;
}

// ---- peer-memory exchange of the sharded loops (replaces the caller's all-gather) --------------------------------
// Every shard owns an exchange region  [parity 2][world][R + 2] doubles + [parity 2][world] 64-bit flags  that all its
// peers can address (NVLink / NVSwitch peer memory; the host passes the peer addresses, e.g. torch symmetric memory).
// k_p2p_push computes this shard's candidate column and STORES it -- header and column -- straight into every peer's
// region (remote st.global over NVLink), then the last CTA releases a generation flag in every region (st.release.sys).
// k_p2p_pull (one warp) acquires the `world` flags of its own region and takes the winner decision.  The compute step
// (gather / replay of the column) and the transfer are one kernel; no collective call sits between the two.
struct P2PPeers {
    double* base[16];   // base[g] = rank g's region as addressed from this GPU
    int32_t world, rank;
    int64_t xstride;    // R + 2
};

__device__ __forceinline__ unsigned long long* p2p_flags(double* region, int world, int64_t xstride) {
    return reinterpret_cast<unsigned long long*>(region + 2 * (int64_t)world * xstride);
}

template <bool LOOKAHEAD>
__global__ void __launch_bounds__(BLK_THREADS)
k_p2p_push(const double* __restrict__ T, int64_t R, int64_t ld, DevState* st, BlkBuffers B, P2PPeers P) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sq[BLK_KMAX];
    __shared__ bool is_last;
    if (st->done) return;
    const int have = st->have_pivot;
    const long long gen = st->xgen + 1;
    const int par = (int)(gen & 1);
    const int64_t slot = ((int64_t)par * P.world + P.rank) * P.xstride;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0) {
        const double v = have ? st->best_val : 0.0, lab = have ? (double)st->enter_lab : -1.0;
        for (int g = 0; g < P.world; ++g) {
            P.base[g][slot] = v;
            P.base[g][slot + 1] = lab;
        }
    }
    if (have) {
        const int s = st->s;
        int t = 0;
        if (LOOKAHEAD) {
            t = (int)(st->n_pivots - B.pend->base);
            if (threadIdx.x < t) {
                sr[threadIdx.x] = B.pend->r[threadIdx.x];
                ss[threadIdx.x] = B.pend->s[threadIdx.x];
                sinv[threadIdx.x] = B.pend->inv_p[threadIdx.x];
                sq[threadIdx.x] = B.qP[(int64_t)threadIdx.x * B.Cpad + s];
            }
            __syncthreads();
        }
        for (int64_t i = tid; i < R; i += nthr) {
            double a = T[i * ld + s];
            if (LOOKAHEAD)
                for (int u = 0; u < t; ++u)
                    a = blk_step(a, i == sr[u], s == ss[u], B.colP[(int64_t)u * B.Rpad + i], sq[u], sinv[u]);
            for (int g = 0; g < P.world; ++g) P.base[g][slot + 2 + i] = a;
        }
    }
    // all stores of this CTA before its ticket; the last CTA publishes the generation in every region
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&st->ticket_push, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) st->ticket_push = 0;
    if ((int)threadIdx.x < P.world) {
        unsigned long long* f = p2p_flags(P.base[threadIdx.x], P.world, P.xstride) + (int64_t)par * P.world + P.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)gen) : "memory");
    }
}

// One warp: wait for the `world` candidates of this generation in the local region, then decide like k_shard_winner.
// A peer that never shows up (crashed rank) must not hang the GPU: after ~2^34 cycles (about 9 s) the loop ends with status 4.
__global__ void k_p2p_pull(DevState* st, P2PPeers P, int bland) {
    if (st->done) return;
    const int lane = threadIdx.x;
    const long long gen = st->xgen + 1;
    const int par = (int)(gen & 1);
    double* region = P.base[P.rank];
    bool ok = true;
    if (lane < P.world) {
        const unsigned long long* f = p2p_flags(region, P.world, P.xstride) + (int64_t)par * P.world + lane;
        const long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v != (unsigned long long)gen && clock64() - t0 > (1ll << 34)) {
                ok = false;
                break;
            }
        } while (v != (unsigned long long)gen);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane != 0) return;
    if (!ok) {
        st->done = 1;
        st->status = 4;
        st->have_pivot = 0;
        return;
    }
    int win = -1;
    double wv = 0.0;
    int32_t wl = B200LP_NO_LAB;
    for (int g = 0; g < P.world; ++g) {
        const double* h = region + ((int64_t)par * P.world + g) * P.xstride;
        const double v = __ldcg(h), labd = __ldcg(h + 1);
        if (labd < 0.0) continue;
        const int32_t lab = (int32_t)labd;
        bool better;
        if (win < 0) better = true;
        else if (bland) better = lab < wl;
        else better = v < wv || (v == wv && lab < wl);
        if (better) {
            win = g;
            wv = v;
            wl = lab;
        }
    }
    st->xgen = gen;
    if (win < 0) {
        st->done = 1;
        st->status = 0;
        st->have_pivot = 0;
        return;
    }
    st->have_pivot = 1;
    st->win_rank = par * P.world + win;  // slot index inside the region: k_ratio / k_blk_ratio read column [win_rank]
    st->enter_lab = wl;
    st->best_val = wv;
    if (win != P.rank) st->s = -1;
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
        for (int u = 0; u < t; ++u)
            a = blk_step(a, i == sr[u], s == ss[u], B.colP[(int64_t)u * B.Rpad + i], sq[u], sinv[u]);
        cand[2 + i] = a;
    }
}

}  // namespace b200lp
