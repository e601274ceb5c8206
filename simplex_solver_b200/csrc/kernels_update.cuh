// kernels_update.cuh -- the pivot: fused row-scale + rank-1 update of the whole tableau, one read and one
// write of every element per pivot (algorithmic traffic 2*R*C*8 bytes; HBM-bound, 1 FMA per 16 bytes).
//
//   q_j     = T[r][j] / p                     (recomputed per tile from the untouched row r)
//   T[i][j] = fma(-col_i, q_j, T[i][j])       i != r
//   T[i][s] = fma(-col_i, 1/p, 0)             i != r   (condensed tableau: column s becomes the column of
//                                                       the leaving variable)
// Row r itself is NOT written here (other CTAs read it); it is scaled by the next k_price / k_flush_row.
// `col` is the contiguous copy of the entering column made by k_ratio.
//
// Two variants of the same arithmetic:
//   k_update_ldg : 128-bit vectorised, coalesced ld.global.cs / st.global.cs, register resident
//   k_update_tma : cp.async.bulk.tensor 2-D tiles into a shared-memory ring (mbarrier full/empty),
//                  in-place FMA in shared memory, cp.async.bulk.tensor store back
//
// Reference seam: the tableau update inside simple_simplex.optimize_json_format
// (/root/reference/app/controllers/solver_controller.py:318).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200lp {

// ------------------------------------------------------------------------------------------------------
// Variant 1: vectorised global loads/stores.  A CTA of NT threads owns tiles of `tile_rows` rows x
// (2*NT) columns; each thread keeps q for its two columns in registers and streams UNROLL rows at a time.
// ------------------------------------------------------------------------------------------------------
template <int NT, int UNROLL>
__global__ void __launch_bounds__(NT)
k_update_ldg(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const double* __restrict__ col,
             const DevState* __restrict__ st, int tile_rows, int tiles_c, int64_t n_tiles) {
    if (st->done || !st->have_pivot) return;
    const int r = st->r, s = st->s;
    const double p = st->p, inv_p = st->inv_p;
    const double* rowr = T + (int64_t)r * ld;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tc = tile % tiles_c, tr = tile / tiles_c;
        const int64_t j = tc * (2 * NT) + 2 * threadIdx.x;
        if (j >= C) continue;
        const double2 rv = *reinterpret_cast<const double2*>(rowr + j);
        const bool sx = (j == s), sy = (j + 1 == s);
        const double qx = sx ? inv_p : rv.x / p;
        const double qy = sy ? inv_p : rv.y / p;
        const int64_t i0 = tr * tile_rows;
        const int64_t i1 = min(R, i0 + (int64_t)tile_rows);
        double* base = T + j;
        for (int64_t i = i0; i < i1; i += UNROLL) {
            double2 t[UNROLL];
            double c[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (i + u < i1) {
                    t[u] = ld_stream(reinterpret_cast<const double2*>(base + (i + u) * ld));
                    c[u] = col[i + u];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (i + u < i1 && i + u != r) {
                    double2 v = t[u];
                    if (sx) v.x = 0.0;
                    if (sy) v.y = 0.0;
                    v.x = __fma_rn(-c[u], qx, v.x);
                    v.y = __fma_rn(-c[u], qy, v.y);
                    st_stream(reinterpret_cast<double2*>(base + (i + u) * ld), v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Variant 2: TMA pipeline.  One producer warp issues cp.async.bulk.tensor.2d loads of BOX_R x BOX_C tiles
// into a STAGES-deep shared-memory ring (mbarrier "full"); 256 consumer threads update the tile in place in
// shared memory (conflict-free 128-bit ld.shared/st.shared); one elected thread stores it back with
// cp.async.bulk.tensor and hands the slot back to the producer (mbarrier "empty") once the store has
// finished READING shared memory (cp.async.bulk.wait_group.read).  Out-of-range parts of edge tiles are
// zero-filled on load and clipped on store by the TMA unit, so ragged R and C need no special code.
// ------------------------------------------------------------------------------------------------------
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t x, int32_t y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void store_2d(const CUtensorMap* map, const void* smem_src, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(smem_src)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma

constexpr int TMA_BOX_C = 256;                   // doubles per tile row (2 KB)
constexpr int TMA_BOX_R = 16;                    // rows per tile -> 32 KB tiles
constexpr int TMA_STAGES = 4;                    // ring depth (4 x 34 KB); deeper rings measured slower on B200
constexpr int TMA_STORE_LAG = 1;                 // stores allowed to be still reading shared memory
constexpr int TMA_CONSUMERS = 256;               // 8 consumer warps
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;  // + 1 producer warp

template <int BOX_R, int STAGES>
struct TmaCfg {
    static constexpr int TILE_DOUBLES = TMA_BOX_C * BOX_R;
    static constexpr int TILE_BYTES = TILE_DOUBLES * 8;
    static constexpr int COL_BYTES = BOX_R * 8;                       // pivot-column slice of the tile's rows
    static constexpr int COL_PAD = ((COL_BYTES + 127) / 128) * 128;
    static constexpr int ROW_BYTES = TMA_BOX_C * 8;                   // pivot-row strip (first tile of a work item)
    static constexpr int STAGE_BYTES = TILE_BYTES + COL_PAD + ROW_BYTES;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 232448, "ring does not fit the 227 KB of shared memory a CTA can opt in to");
    // STORE_LAG = stores allowed to be still reading shared memory when the next tile is processed
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tma::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tma::smem_u32(bar))
                 : "memory");
}

// Work item = (column strip of TMA_BOX_C columns) x (chunk of `chunk_tiles` row tiles); a CTA walks its items
// with a grid stride, one resident CTA per SM.  Each stage carries the tile, the slice of the pivot column for
// the tile's rows and -- on the first tile of a work item -- the pivot-row strip (1-D bulk copies on the same
// mbarrier), so the consumers never wait on a global load: everything they read arrives through the TMA unit.
template <int BOX_R, int STAGES, int STORE_LAG, int CTAS_PER_SM>
__global__ void __launch_bounds__(TMA_THREADS, CTAS_PER_SM)
k_update_tma(const __grid_constant__ CUtensorMap map, const double* __restrict__ T, int64_t R, int64_t C, int64_t ld,
             const double* __restrict__ col, const DevState* __restrict__ st, int strips, int chunk_tiles,
             int64_t n_work) {
    using Cfg = TmaCfg<BOX_R, STAGES>;
    if (st->done || !st->have_pivot) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;

    const int r = st->r, s = st->s;
    const double p = st->p, inv_p = st->inv_p;
    const int warp = threadIdx.x >> 5;
    const int64_t row_tiles = (R + BOX_R - 1) / BOX_R;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp == TMA_CONSUMERS / 32) {
        // ===== producer warp: one elected lane keeps the ring full =====
        if ((threadIdx.x & 31) == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int64_t strip = w % strips, chunk = w / strips;
                const int64_t rt0 = chunk * chunk_tiles, rt1 = min(row_tiles, rt0 + (int64_t)chunk_tiles);
                for (int64_t rt = rt0; rt < rt1; ++rt) {
                    uint8_t* slot = smem + (size_t)stage * Cfg::STAGE_BYTES;
                    tma::mbar_wait(&empty[stage], phase ^ 1);
                    const uint32_t row_bytes = (rt == rt0) ? (uint32_t)(min((int64_t)TMA_BOX_C, ld - strip * TMA_BOX_C) * 8) : 0u;
                    tma::mbar_expect_tx(&full[stage], Cfg::TILE_BYTES + Cfg::COL_BYTES + row_bytes);
                    tma::load_2d(slot, &map, &full[stage], (int32_t)(strip * TMA_BOX_C), (int32_t)(rt * BOX_R));
                    bulk_load_1d(slot + Cfg::TILE_BYTES, col + rt * BOX_R, Cfg::COL_BYTES, &full[stage]);
                    if (row_bytes)
                        bulk_load_1d(slot + Cfg::TILE_BYTES + Cfg::COL_PAD, T + (int64_t)r * ld + strip * TMA_BOX_C,
                                     row_bytes, &full[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        return;
    }

    // ===== consumers: thread t owns the column pair 2*(t%128) of the strip and the rows of parity t/128,
    //       so a warp touches 512 contiguous bytes of a tile row: conflict-free 128-bit accesses =====
    const int t = threadIdx.x;
    const int cp = (t & 127) * 2;
    const int rpar = t >> 7;
    int stage = 0;
    uint32_t phase = 0;
    long long issued = 0, released = 0;  // used by thread 0 only
    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int64_t strip = w % strips, chunk = w / strips;
        const int64_t rt0 = chunk * chunk_tiles, rt1 = min(row_tiles, rt0 + (int64_t)chunk_tiles);
        const int64_t j = strip * TMA_BOX_C + cp;
        const bool sx = (j == s), sy = (j + 1 == s);
        double qx = 0.0, qy = 0.0;
        for (int64_t rt = rt0; rt < rt1; ++rt) {
            const int64_t i0 = rt * BOX_R;
            uint8_t* slot = smem + (size_t)stage * Cfg::STAGE_BYTES;
            double* tile = reinterpret_cast<double*>(slot);
            const double* cs = reinterpret_cast<const double*>(slot + Cfg::TILE_BYTES);
            tma::mbar_wait(&full[stage], phase);
            if (rt == rt0 && j < ld) {
                const double2 rv = *reinterpret_cast<const double2*>(slot + Cfg::TILE_BYTES + Cfg::COL_PAD + cp * 8);
                qx = sx ? inv_p : rv.x / p;
                qy = sy ? inv_p : rv.y / p;
            }
#pragma unroll
            for (int k = 0; k < BOX_R / 2; ++k) {
                const int lr = 2 * k + rpar;
                if (i0 + lr != r) {
                    const double c = cs[lr];
                    double2* cell = reinterpret_cast<double2*>(tile + lr * TMA_BOX_C + cp);
                    double2 v = *cell;
                    if (sx) v.x = 0.0;
                    if (sy) v.y = 0.0;
                    v.x = __fma_rn(-c, qx, v.x);
                    v.y = __fma_rn(-c, qy, v.y);
                    *cell = v;
                }
            }
            tma::fence_proxy_async();  // generic-proxy writes -> visible to the async proxy (TMA store)
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
            if (t == 0) {
                tma::store_2d(&map, tile, (int32_t)(strip * TMA_BOX_C), (int32_t)i0);
                tma::store_commit();
                ++issued;
                tma::store_wait_read<STORE_LAG>();
                while (released < issued - STORE_LAG) {
                    tma::mbar_arrive(&empty[released % STAGES]);
                    ++released;
                }
            }
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
    }
    if (t == 0) tma::store_wait_all();
}

}  // namespace b200lp
