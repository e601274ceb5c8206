// kernels_shard.cuh -- one pivot decision of a COLUMN-SHARDED tableau in ONE kernel per GPU: local pricing, the
// candidate exchange over NVLink peer memory, the winner decision and the ratio test (BASELINE config 5).
//
// It replaces the launch chain  k_price<SHARDED> -> k_p2p_push -> k_p2p_pull -> k_ratio  (four kernels whose tails and
// launch gaps were serial on every pivot: ~0.13 ms of a 5.6 ms pivot at 8 GPUs) and, in the look-ahead loop,
// k_blk_rowprice<SHARDED> -> k_p2p_push<true> -> k_p2p_pull -> k_blk_ratio.  One thread-block cluster of 16 CTAs per GPU:
//
//   phase A (own columns) : [rank-1: finish the previous pivot's row scaling | look-ahead: row part of the previous
//                           pivot] + argmin over the own slice of the objective row -> cluster reduction (DSMEM)
//   phase X (exchange)    : the candidate record [reduced cost, variable id, column (R doubles)] is STORED by all 16 CTAs
//                           straight into slot `rank` of every peer's exchange region (st.global on NVLink-mapped peer
//                           pointers, 16 bytes per store); fence.sys + cluster barrier; one st.release.sys per peer
//                           publishes the generation; one warp per CTA polls the `world` generation words of the LOCAL
//                           region with ld.acquire.sys; every CTA takes the same winner from the `world` headers
//   phase B (all rows)    : ratio test on the winner's column (read from the local region, contiguous) against the RHS
//                           replica -> cluster reduction -> bookkeeping by CTA 0 (labels, history, DevState)
//
// followed by the ordinary update kernel (rank-1 loop) or nothing (look-ahead loop: the flush comes every K pivots).
//
// Why every shard pushes its column instead of "argmin exchange, then the winner broadcasts": the winner is only known
// after a first exchange, so header-first puts TWO NVLink round trips on the critical path of a pivot, and the column
// that then crosses the links (R doubles to world - 1 peers from ONE GPU) takes exactly as long as world GPUs pushing
// theirs concurrently through the switch (each GPU's egress and ingress carry (world - 1) * 8R bytes either way: 7 MiB =
// 9.5 us at 770 GB/s for R = 131072, world = 8).  All-push costs bytes that the links have to spare (0.02 % of a pivot's
// HBM bytes) and saves a round trip on the chain that is serial.
//
// Two region halves alternate by generation parity: a rank needs everybody's generation g before it can push g + 1, so it
// is never more than one exchange ahead of the slowest reader.  The generation counter lives outside DevState (which is
// cleared by attach / generate): flags left over from earlier runs must never equal a future generation.
//
// Single-GPU emulation (tests): a launch takes an ARRAY of shard contexts and runs one cluster per context, all resident
// at once (cooperative launch), so shards that wait for one another are never separate launches (which nothing would
// force to run concurrently).  Production: one context, one cluster.
#pragma once
#include "kernels_cluster.cuh"

namespace b200lp {

struct ShardCtx {
    PickArgs A;         // tableau binding, labels, state, history, look-ahead buffers of THIS shard (obj_row, eps unused)
    P2PPeers P;         // exchange regions of all shards as addressed from this GPU
    long long* xgen;    // generation of the last completed exchange (persistent: never reset while connected)
    int32_t* error;     // set to 1 when a peer's generation never arrives
};

constexpr int SHARD_CLUSTER = 16;

// PRE (look-ahead loop only): the candidate record of this exchange has already been pushed to every peer, generation
// word included, by k_blk_shard_push (below) right before this launch; the kernel only prices the current objective row
// (to know its own candidate's position), waits for the peers' records, decides and runs the ratio test.
template <bool BLAND, bool BLOCKED, bool PRE = false>
__global__ void __launch_bounds__(CL_THREADS, 1)
k_shard_pick(const ShardCtx* __restrict__ ctxs, int64_t obj_row, double eps_cost, double eps_pivot) {
    static_assert(!PRE || BLOCKED, "the pre-pushed exchange belongs to the look-ahead loop");
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ Key sk[CL_THREADS / 32];
    __shared__ Key slot_price, slot_ratio, bc;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sx[BLK_KMAX];
    __shared__ int win_s;      // winner's slot (rank) of this exchange, -1 none, -2 a peer went missing
    __shared__ Key win_k;
    const ShardCtx& X = ctxs[blockIdx.x / cluster.num_blocks()];
    const PickArgs& A = X.A;
    const P2PPeers& P = X.P;
    DevState* st = A.st;
    if (st->done) return;  // uniform over all shards: every shard takes the same decisions
    const int64_t gtid = (int64_t)cluster.block_rank() * CL_THREADS + threadIdx.x;
    const int64_t nthr = (int64_t)cluster.num_blocks() * CL_THREADS;
    const int64_t R = A.R, C = A.C, ld = A.ld;
    const long long n_piv = st->n_pivots;
    const long long base = BLOCKED ? A.B.pend->base : 0;
    const bool leader = cluster.block_rank() == 0 && threadIdx.x == 0;

    // ------------------------------------------------ phase A: own columns -------------------------------------------
    Key k = key_none();
    if (BLOCKED) {
        if (st->pend != 0) {  // row part of the previous look-ahead pivot (as k_pick_cluster / k_blk_rowprice)
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
            const double c_obj = colT[obj_row];
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
                    if (lab < A.art_base && d < -eps_cost) {
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
            __syncthreads();  // sr/ss/sinv/sx are reloaded below
        } else {
            for (int64_t j = gtid; j < C - 1; j += nthr) {
                const int32_t lab = A.collab[j];
                const double d = A.B.objcur[j];
                if (lab < A.art_base && d < -eps_cost) {
                    Key c;
                    c.v = d;
                    c.lab = lab;
                    c.pos = (int32_t)j;
                    k = key_min<BLAND>(k, c);
                }
            }
        }
    } else {
        if (st->pend) {  // deferred scaling of the previous pivot row on the own slice (s = -1: the column lives elsewhere)
            const int r = st->r, s = st->s;
            const double p = st->p, inv_p = st->inv_p;
            double* row = A.T + (int64_t)r * ld;
            for (int64_t j = gtid; j < C; j += nthr) {
                const double v = row[j];
                row[j] = (j == s) ? inv_p : v / p;
            }
        }
        const double* d = A.T + obj_row * ld;
        for (int64_t j = gtid; j < C - 1; j += nthr) {
            const int32_t lab = A.collab[j];
            const double v = d[j];
            if (lab < A.art_base && v < -eps_cost) {
                Key c;
                c.v = v;
                c.lab = lab;
                c.pos = (int32_t)j;
                k = key_min<BLAND>(k, c);
            }
        }
    }
    const Key mine = cluster_key_min<BLAND>(cluster, k, &slot_price, sk, &bc);
    if (n_piv >= st->max_pivots) {  // replicated counter: every shard stops here, before touching the exchange
        cluster.sync();
        if (leader) {
            st->pend = 0;
            st->have_pivot = 0;
            st->done = 1;
            st->status = 1;
        }
        return;
    }

    // ------------------------------------------------ phase X: exchange ----------------------------------------------
    const int world = P.world, rank = P.rank;
    const long long gen = *X.xgen + 1;
    const int par = (int)(gen & 1);
    const int64_t slot = ((int64_t)par * world + rank) * P.xstride;
    const bool have = mine.lab != B200LP_NO_LAB;
    int t = 0;
    if (BLOCKED) t = (int)(n_piv - base);
    if (!PRE) {
    if (leader) {
        const double2 h = make_double2(have ? mine.v : 0.0, have ? (double)mine.lab : -1.0);
        for (int g = 0; g < world; ++g) *reinterpret_cast<double2*>(P.base[g] + slot) = h;  // xstride is even
    }
    if (have) {
        const int s = mine.pos;
        if (BLOCKED) {
            if (threadIdx.x < t) {
                sr[threadIdx.x] = A.B.pend->r[threadIdx.x];
                ss[threadIdx.x] = A.B.pend->s[threadIdx.x];
                sinv[threadIdx.x] = A.B.pend->inv_p[threadIdx.x];
                sx[threadIdx.x] = __ldcg(A.B.qP + (int64_t)threadIdx.x * A.B.Cpad + s);  // q_u[s]; q_{t-1} is fresh from phase A
            }
            __syncthreads();
        }
        // two rows per thread: one 16-byte store per peer (the record's column starts at an even offset)
        for (int64_t i = 2 * gtid; i < R; i += 2 * nthr) {
            double a0 = __ldcg(A.T + i * ld + s);
            double a1 = (i + 1 < R) ? __ldcg(A.T + (i + 1) * ld + s) : 0.0;
            if (BLOCKED) {
                a0 = blk_replay<false, PICK_BATCH>(a0, t, A.B.colP + i, A.B.Rpad, sx, sr, ss, sinv, i, s);
                if (i + 1 < R) a1 = blk_replay<false, PICK_BATCH>(a1, t, A.B.colP + i + 1, A.B.Rpad, sx, sr, ss, sinv, i + 1, s);
            }
            const double2 v = make_double2(a0, a1);
            for (int g = 0; g < world; ++g) *reinterpret_cast<double2*>(P.base[g] + slot + 2 + i) = v;
        }
    }
    __threadfence_system();  // this thread's remote stores before the cluster barrier that precedes the flags
    cluster.sync();
    if (cluster.block_rank() == 0 && (int)threadIdx.x < world) {
        unsigned long long* f = p2p_flags(P.base[threadIdx.x], world, P.xstride) + (int64_t)par * world + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)gen) : "memory");
    }
    }
    // every CTA waits for the `world` generations of the local region and takes the same decision
    double* region = P.base[rank];
    if (threadIdx.x < 32) {
        bool ok = true;
        Key c = key_none();
        if ((int)threadIdx.x < world) {
            const unsigned long long* f = p2p_flags(region, world, P.xstride) + (int64_t)par * world + threadIdx.x;
            unsigned long long v;
            long long t0 = 0;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v != (unsigned long long)gen) {  // a peer that never shows up must not hang the GPU (~4 s)
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > (1ll << 33)) {
                        ok = false;
                        break;
                    }
                }
            } while (v != (unsigned long long)gen);
            if (ok) {
                const double2 h = __ldcg(reinterpret_cast<const double2*>(region + ((int64_t)par * world + threadIdx.x) * P.xstride));
                if (h.y >= 0.0) {
                    c.v = h.x;
                    c.lab = (int32_t)h.y;
                    c.pos = (int32_t)threadIdx.x;
                }
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        c = warp_key_min<BLAND>(c);
        if (threadIdx.x == 0) {
            win_k = c;
            win_s = !ok ? -2 : (c.lab == B200LP_NO_LAB ? -1 : c.pos);
        }
    }
    __syncthreads();
    const int wslot = win_s;
    const Key win = win_k;
    if (wslot < 0) {
        cluster.sync();
        if (leader) {
            *X.xgen = gen;
            st->pend = 0;
            st->have_pivot = 0;
            st->done = 1;
            st->status = wslot == -2 ? 4 : 0;  // a missing peer is an error; no candidate anywhere: OPTIMAL
            st->s = -1;
            st->enter_lab = -1;
            if (wslot == -2) *X.error = 1;
        }
        return;
    }

    // ------------------------------------------------ phase B: all rows ----------------------------------------------
    const double* wcol = region + ((int64_t)par * world + wslot) * P.xstride + 2;
    double* colT = BLOCKED ? A.B.colP + (int64_t)t * A.B.Rpad : A.col;
    k = key_none();
    for (int64_t i = gtid; i < R; i += nthr) {
        const double a = __ldcg(wcol + i);
        const double rhs = BLOCKED ? __ldcg(A.B.rhscur + i) : __ldcg(A.T + i * ld + C - 1);
        colT[i] = a;
        if (i < A.m) {
            const int32_t lab = A.rowlab[i];
            if (lab >= 0 && a > eps_pivot) {
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
    *X.xgen = gen;
    const int r = lw.pos;
    if (r < 0) {
        st->pend = 0;
        st->done = 1;
        st->status = 3;  // UNBOUNDED
        st->have_pivot = 0;
        return;
    }
    const int s = (wslot == rank) ? mine.pos : -1;  // position in THIS shard, -1 when the column lives in another one
    const double p = __ldcg(wcol + r);
    const double inv_p = 1.0 / p;
    st->s = s;
    st->enter_lab = win.lab;
    st->best_val = win.v;
    st->win_rank = par * world + wslot;
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
    if (s >= 0) A.collab[s] = leave;
    if (n_piv < A.hist_cap) {
        A.h_row[n_piv] = r;
        A.h_col[n_piv] = s;
        A.h_enter[n_piv] = win.lab;
        A.h_leave[n_piv] = leave;
    }
    st->n_pivots = n_piv + 1;
}

// Look-ahead loop, first half of the exchange on ALL SMs: the candidate column of this shard (chosen by
// k_blk_rowprice<.., SHARDED>: DevState.have_pivot / s / enter_lab / best_val) is gathered, brought up to date by replaying
// the pending steps and stored straight into slot `rank` of every peer's region -- the work k_shard_pick<.., BLOCKED> does
// on one 16-CTA cluster, where the t * R * 8 bytes of history and the (world - 1) * R * 8 bytes of NVLink stores pass
// through one GPC (52-115 us per decision on a 131072-row shard against ~15 us here).  The last CTA to finish (device
// ticket, after a system-scope fence by every storing thread) publishes the generation word to every peer with
// st.release.sys.  k_shard_pick<.., true, true> follows and completes the exchange.
__global__ void __launch_bounds__(BLK_THREADS)
k_blk_shard_push(const ShardCtx* __restrict__ ctxs, int ctx_index) {
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    __shared__ double sinv[BLK_KMAX], sq[BLK_KMAX];
    __shared__ bool is_last;
    const ShardCtx& X = ctxs[ctx_index];
    const PickArgs& A = X.A;
    const P2PPeers& P = X.P;
    DevState* st = A.st;
    if (st->done) return;
    const int world = P.world, rank = P.rank;
    const long long gen = *X.xgen + 1;
    const int par = (int)(gen & 1);
    const int64_t slot = ((int64_t)par * world + rank) * P.xstride;
    const bool have = st->have_pivot != 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0) {
        const double2 h = make_double2(have ? st->best_val : 0.0, have ? (double)st->enter_lab : -1.0);
        for (int g = 0; g < world; ++g) *reinterpret_cast<double2*>(P.base[g] + slot) = h;
    }
    if (have) {
        const int s = st->s;
        const int t = (int)(st->n_pivots - A.B.pend->base);
        if (threadIdx.x < t) {
            sr[threadIdx.x] = A.B.pend->r[threadIdx.x];
            ss[threadIdx.x] = A.B.pend->s[threadIdx.x];
            sinv[threadIdx.x] = A.B.pend->inv_p[threadIdx.x];
            sq[threadIdx.x] = A.B.qP[(int64_t)threadIdx.x * A.B.Cpad + s];
        }
        __syncthreads();
        const int64_t R = A.R, ld = A.ld;
        for (int64_t i = 2 * tid; i < R; i += 2 * nthr) {
            double a0 = A.T[i * ld + s];
            double a1 = (i + 1 < R) ? A.T[(i + 1) * ld + s] : 0.0;
            a0 = blk_replay<false>(a0, t, A.B.colP + i, A.B.Rpad, sq, sr, ss, sinv, i, s);
            if (i + 1 < R) a1 = blk_replay<false>(a1, t, A.B.colP + i + 1, A.B.Rpad, sq, sr, ss, sinv, i + 1, s);
            const double2 v = make_double2(a0, a1);
            for (int g = 0; g < world; ++g) *reinterpret_cast<double2*>(P.base[g] + slot + 2 + i) = v;
        }
    }
    __threadfence_system();  // this thread's remote stores are performed before its CTA takes a ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&st->ticket_push, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < world) {
        unsigned long long* f = p2p_flags(P.base[threadIdx.x], world, P.xstride) + (int64_t)par * world + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)gen) : "memory");
    }
    if (threadIdx.x == 0) st->ticket_push = 0;
}

}  // namespace b200lp
