// kernels_pick.cuh -- entering-variable selection (argmin over reduced costs) and the masked min-ratio
// test, each a multi-CTA reduction finished by the last CTA to arrive ("ticket" pattern), so that one
// launch leaves the decision in the device-resident DevState and the host never sees it.
//
// Reference seam: these are the selection steps of the pivot loop behind
// simple_simplex.optimize_json_format (/root/reference/app/controllers/solver_controller.py:318); the rules
// (Dantzig / Bland, lowest-id tie-break) are the ones BASELINE.json's north_star names.
#pragma once
#include "common.cuh"

namespace b200lp {

// Multi-CTA reductions finished by the last CTA to arrive.  (One fat CTA per kernel was measured slower at
// 16384 x 16384: the strided column gather of the ratio test is limited by the loads ONE SM can keep in flight --
// 40 us against 7 us with 64 CTAs.)  A single-CTA launch (small tableaux) skips partials, fence and ticket.
constexpr int PICK_THREADS = 256;

// Scale the row of the previous pivot (deferred so that the update kernel never writes row r while other
// CTAs still read it):  T[r][j] = T[r][j] / p,  T[r][s] = 1/p.
__device__ __forceinline__ void flush_row_slice(double* __restrict__ T, int64_t C, int64_t ld, const DevState* st,
                                                int64_t j0, int64_t j1, int64_t stride) {
    const int r = st->r, s = st->s;
    const double p = st->p, inv_p = st->inv_p;
    double* row = T + (int64_t)r * ld;
    for (int64_t j = j0; j < j1; j += stride) {
        double v = row[j];
        row[j] = (j == s) ? inv_p : v / p;
    }
}

__global__ void __launch_bounds__(PICK_THREADS) k_flush_row(double* T, int64_t C, int64_t ld, DevState* st) {
    if (!st->pend) return;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    flush_row_slice(T, C, ld, st, tid, C, (int64_t)gridDim.x * blockDim.x);
    // the flag is cleared by k_clear_pend (next launch on the stream) so that every CTA above saw it set
}
__global__ void k_clear_pend(DevState* st) { st->pend = 0; }

// Phase A of an iteration.  (1) finish the previous pivot's row scaling, (2) argmin over the objective row.
// SHARDED: "no local candidate" is not the end of the loop (another shard may have one); the decision is
// taken by k_shard_winner after the all-gather.
template <bool BLAND, bool SHARDED>
__global__ void __launch_bounds__(PICK_THREADS)
k_price(double* T, int64_t C, int64_t ld, int64_t obj_row, const int32_t* __restrict__ collab, int32_t art_base,
        double eps_cost, DevState* st, Key* partials) {
    __shared__ Key sk[PICK_THREADS / 32];
    __shared__ bool is_last;
    const int done = st->done, pend = st->pend;
    if (done && !pend) return;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (pend) flush_row_slice(T, C, ld, st, tid, C, nthr);

    Key k = key_none();
    if (!done) {
        const double* d = T + obj_row * ld;
        for (int64_t j = tid; j < C - 1; j += nthr) {
            const int32_t lab = collab[j];
            const double v = d[j];
            if (lab < art_base && v < -eps_cost) {
                Key c;
                c.v = v;
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
            const unsigned int t = atomicAdd(&st->ticket_price, 1u);
            is_last = (t == gridDim.x - 1);
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
        st->pend = 0;
        if (!done) {
            if (st->n_pivots >= st->max_pivots) {
                st->done = 1;
                st->status = 1;  // LIMIT
                st->have_pivot = 0;
            } else if (k.lab == B200LP_NO_LAB) {
                st->have_pivot = 0;
                st->s = -1;
                st->enter_lab = -1;
                if (!SHARDED) {
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
}

// Sharded runs: publish this shard's candidate [reduced cost, variable id, column entries] for the all-gather.
__global__ void __launch_bounds__(PICK_THREADS)
k_shard_extract(const double* __restrict__ T, int64_t R, int64_t ld, const DevState* st, double* __restrict__ cand) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int have = st->have_pivot && !st->done;
    if (tid == 0) {
        cand[0] = have ? st->best_val : 0.0;
        cand[1] = have ? (double)st->enter_lab : -1.0;
    }
    if (!have) return;
    const int s = st->s;
    for (int64_t i = tid; i < R; i += (int64_t)gridDim.x * blockDim.x) cand[2 + i] = T[i * ld + s];
}

// Sharded runs: every shard takes the same decision from the gathered headers (total order => identical).
__global__ void k_shard_winner(const double* __restrict__ gathered, int64_t stride, int world, int rank, int bland,
                               DevState* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st->done) return;
    int win = -1;
    double wv = 0.0;
    int32_t wl = B200LP_NO_LAB;
    for (int g = 0; g < world; ++g) {
        const double v = gathered[g * stride];
        const double labd = gathered[g * stride + 1];
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
    if (win < 0) {
        st->done = 1;
        st->status = 0;
        st->have_pivot = 0;
        return;
    }
    st->have_pivot = 1;
    st->win_rank = win;
    st->enter_lab = wl;
    st->best_val = wv;
    if (win != rank) st->s = -1;
}

// Phase B of an iteration: copy the entering column into `col` (contiguous, read by the update kernel) and
// run the masked min-ratio reduction over it.  ROW_PRESET: the row is already in st->r (drive-out pivots and
// b200lp_pivot), only the column copy and the bookkeeping are done.
template <bool ROW_PRESET>
__global__ void __launch_bounds__(PICK_THREADS)
k_ratio(const double* __restrict__ T, int64_t R, int64_t m, int64_t C, int64_t ld, int32_t* rowlab, int32_t* collab,
        double eps_pivot, DevState* st, Key* partials, double* __restrict__ col, const double* __restrict__ ext,
        int64_t ext_stride, int32_t* h_row, int32_t* h_col, int32_t* h_enter, int32_t* h_leave, int64_t hist_cap) {
    __shared__ Key sk[PICK_THREADS / 32];
    __shared__ bool is_last;
    if (st->done || !st->have_pivot) return;
    const int s = st->s;
    const double* src = ext ? ext + (int64_t)st->win_rank * ext_stride + 2 : nullptr;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    Key k = key_none();
    for (int64_t i = tid; i < R; i += nthr) {
        const double a = src ? src[i] : T[i * ld + s];
        col[i] = a;
        if (!ROW_PRESET && i < m) {
            const int32_t lab = rowlab[i];
            if (lab >= 0 && a > eps_pivot) {
                Key c;
                c.v = T[i * ld + C - 1] / a;
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
            const unsigned int t = atomicAdd(&st->ticket_ratio, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        k = key_none();
        if (!ROW_PRESET) {
            for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
                Key c;
                c.v = __ldcg(&partials[b].v);
                c.lab = __ldcg(&partials[b].lab);
                c.pos = __ldcg(&partials[b].pos);
                k = key_min<false>(k, c);
            }
        }
        __syncthreads();
        k = block_key_min<false>(k, sk);
    }
    if (threadIdx.x == 0) {
        st->ticket_ratio = 0;
        int r;
        if (ROW_PRESET) r = st->r;
        else r = k.pos;
        if (r < 0) {
            st->done = 1;
            st->status = 3;  // UNBOUNDED
            st->have_pivot = 0;
        } else {
            const double p = src ? src[r] : T[(int64_t)r * ld + s];
            st->r = r;
            st->p = p;
            st->inv_p = 1.0 / p;
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
            st->pend = 1;
        }
    }
}

// After phase 1: choose the next artificial still basic (lowest row) and the eligible entry of largest
// magnitude in its row (ties to the lowest variable id); rows without one are flagged redundant.
// One CTA; the row scan is O(m), the column scan O(C).
__global__ void __launch_bounds__(1024)
k_driveout_pick(const double* __restrict__ T, int64_t m, int64_t C, int64_t ld, int32_t* rowlab,
                const int32_t* __restrict__ collab, int32_t art_base, double eps_pivot, DevState* st) {
    __shared__ Key sk[32];
    __shared__ Key bc;
    if (st->done) return;
    for (;;) {
        Key k = key_none();
        for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
            if (rowlab[i] >= art_base) {
                Key c;
                c.v = 0.0;
                c.lab = (int32_t)i;
                c.pos = (int32_t)i;
                k = key_min<true>(k, c);
            }
        }
        k = block_key_min<true>(k, sk);
        if (threadIdx.x == 0) bc = k;
        __syncthreads();
        const int row = bc.pos;
        __syncthreads();
        if (row < 0) {
            if (threadIdx.x == 0) {
                st->done = 1;
                st->status = 0;
                st->have_pivot = 0;
            }
            return;
        }
        const double* tr = T + (int64_t)row * ld;
        k = key_none();
        for (int64_t j = threadIdx.x; j < C - 1; j += blockDim.x) {
            const int32_t lab = collab[j];
            const double a = fabs(tr[j]);
            if (lab < art_base && a > eps_pivot) {
                Key c;
                c.v = -a;
                c.lab = lab;
                c.pos = (int32_t)j;
                k = key_min<false>(k, c);
            }
        }
        k = block_key_min<false>(k, sk);
        if (threadIdx.x == 0) bc = k;
        __syncthreads();
        const Key best = bc;
        __syncthreads();
        if (best.lab == B200LP_NO_LAB) {
            if (threadIdx.x == 0) rowlab[row] = -1 - rowlab[row];
            __syncthreads();
            continue;
        }
        if (threadIdx.x == 0) {
            if (st->n_pivots >= st->max_pivots) {
                st->done = 1;
                st->status = 1;
                st->have_pivot = 0;
            } else {
                st->have_pivot = 1;
                st->r = row;
                st->s = best.pos;
                st->enter_lab = best.lab;
            }
        }
        return;
    }
}

// Optional (small LPs that feed the reference's "pivotSteps", solver_controller.py:332-362): keep a dense copy of
// the tableau after every pivot.  Row r is still unscaled at this point, so the scaling is applied on the fly.
__global__ void __launch_bounds__(256)
k_snapshot(const double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* __restrict__ st,
           double* __restrict__ snaps, int64_t cap) {
    if (st->done || !st->have_pivot) return;
    const long long slot = st->n_pivots - 1;
    if (slot < 0 || slot >= cap) return;
    const int r = st->r, s = st->s;
    const double p = st->p, inv_p = st->inv_p;
    double* out = snaps + slot * R * C;
    const int64_t total = R * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / C, j = e - i * C;
        double v = T[i * ld + j];
        if (i == r) v = (j == s) ? inv_p : v / p;
        out[e] = v;
    }
}

__global__ void k_set_pivot(DevState* st, int32_t r, int32_t s, const int32_t* collab) {
    st->done = 0;
    st->have_pivot = 1;
    st->r = r;
    st->s = s;
    st->enter_lab = collab[s];
    st->win_rank = 0;
}

__global__ void k_reset_state(DevState* st, long long max_pivots, int keep_count) {
    st->done = 0;
    st->status = 0;
    st->have_pivot = 0;
    st->s = -1;
    st->r = -1;
    st->win_rank = 0;
    st->ticket_price = 0;
    st->ticket_ratio = 0;
    st->max_pivots = max_pivots;
    if (!keep_count) {
        st->n_pivots = 0;
        st->pend = 0;
    }
}

}  // namespace b200lp
