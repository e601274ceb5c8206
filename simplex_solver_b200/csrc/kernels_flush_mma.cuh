// kernels_flush_mma.cuh -- the main kernel of the look-ahead flush on the FP64 tensor pipe (DMMA).
//
// The flush applies, to every element outside the pivot rows / pivot column pairs of the pending steps, the chain
//     v <- fma(-col_u[i], q_u[j], v),   u = 0 .. t-1 (in this order, one rounding per step).
// That is a rank-t update T <- T - colP^T * qP, but not one a GEMM library may compute: the order of the t roundings is
// part of the arithmetic contract (DESIGN.md section 2 -- every stored double is bit-identical to the sequential loop).
// mma.sync.m8n8k4.f64 on sm_100a evaluates each output element as exactly that chain, k = 0 first:
//     d = fma(a3, b3, fma(a2, b2, fma(a1, b1, fma(a0, b0, c))))
// (scripts/probe_dmma.cu: 12.6 M random elements over exponent spans +-0 .. +-300 equal to the chain in every bit, and
// different from the reverse order and from a pairwise sum; the parity tests of the look-ahead loop are the standing
// check).  So four pending steps of an 8 x 8 tile are ONE instruction whose operands are 1 + 1 doubles per thread:
//
//   k_blk_flush_db (DFMA): per thread and step 8 col_u + 4 q_u doubles from shared memory for 32 FMAs -- the shared-memory
//                          pipe (128 B/clk per SM) is the bottleneck: 73 % busy at 56 % of the FP64 pipe, 0.92 ms per
//                          block on 16384^2 (0.71 of the HBM roofline);
//   this kernel     (DMMA): a warp owns 32 x 32 elements = 4 x 4 tiles; per 4 steps it loads 4 + 4 operand doubles per
//                          thread (16 wavefronts) for 16 DMMAs = 4096 FMAs: 6 x less shared-memory traffic, and the DMMA
//                          pipe is separate from the DFMA pipe (probe: mixed warps take max, not sum).
//
// Layout: CTA = 8 warps side by side, strip = 256 columns (2 KB of every row), tile = 128 rows; a CTA walks DOWN a strip
// (contiguous range of tiles in strip-major order, as k_blk_flush_db), so -q_u of the strip is staged once per strip;
// the tile's col_u slices arrive by cp.async one tile ahead (double buffer); the C fragments of the NEXT 32-row group are
// loaded into registers right after the first k-block of the current one (same scoreboard discipline as the DFMA kernel).
// Fragment layout of m8n8k4.f64: A[i][k] in lane 4 i + k, B[k][j] in lane 4 j + k, C[i][2 c], C[i][2 c + 1] in lane 4 i + c
// -- a thread's C fragment is one 16-byte global access.  Shared-memory rows are padded by 8 doubles so that the 32 lanes
// of an operand load (4 steps x 8 rows/columns) hit 32 distinct 8-byte words of a 256-byte window.
#pragma once
#include "kernels_blocked.cuh"

namespace b200lp {

constexpr int FM_TR = 128;             // rows per tile
constexpr int FM_SW = 256;             // columns per strip (8 warps x 32)
constexpr int FM_PAD = 8;              // padding of a shared-memory row (doubles)
constexpr int FM_KP = BLK_KMAX;        // steps, padded to a multiple of 4 with zero operands
constexpr size_t FM_SMEM_BYTES = (size_t)FM_KP * ((FM_SW + FM_PAD) + 2 * (FM_TR + FM_PAD)) * 8;
static_assert(FM_SMEM_BYTES <= 227 * 1024, "flush shared memory");
static_assert(BLK_KMAX % 4 == 0, "steps are consumed four at a time");

__device__ __forceinline__ void dmma_m8n8k4(double2& d, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d.x), "+d"(d.y)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
k_blk_flush_mma(double* __restrict__ T, int64_t R, int64_t C, int64_t ld, const DevState* st, BlkBuffers B, int64_t n_rb,
                int64_t n_strips) {
    constexpr int QP = FM_SW + FM_PAD, CP = FM_TR + FM_PAD, NG = FM_TR / 32;
    __shared__ int32_t sr[BLK_KMAX], ss[BLK_KMAX];
    extern __shared__ __align__(16) double dyn_fm[];
    double* sq = dyn_fm;                   // [t4][QP]       -q_u of the strip
    double* scol = dyn_fm + FM_KP * QP;    // [2][t4][CP]    col_u of the tile (double buffer)
    const int t = (int)(st->n_pivots - B.pend->base);
    if (t == 0) return;
    const int t4 = (t + 3) & ~3;
    if (threadIdx.x < t) {
        sr[threadIdx.x] = B.pend->r[threadIdx.x];
        ss[threadIdx.x] = B.pend->s[threadIdx.x];
    }
    // zero operands for the steps that pad t to a multiple of 4: fma(0, 0, v) = v
    for (int e = threadIdx.x; e < (t4 - t) * QP; e += 256) sq[t * QP + e] = 0.0;
    for (int e = threadIdx.x; e < (t4 - t) * CP; e += 256) {
        scol[t * CP + e] = 0.0;
        scol[FM_KP * CP + t * CP + e] = 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int li = lane >> 2, lk = lane & 3;  // fragment coordinates: row / column li of a tile, step lk of a k-block
    const int64_t n_items = n_rb * n_strips;
    const int64_t it0 = n_items * blockIdx.x / gridDim.x, it1 = n_items * (blockIdx.x + 1) / gridDim.x;
    if (it0 >= it1) return;

    auto stage_cols = [&](int64_t rb, int buf) {
        const int64_t i0 = rb * FM_TR;
        const int rows = (int)(min(R, i0 + (int64_t)FM_TR) - i0);
        const int pieces = (rows + 1) / 2;  // colP is padded to Rpad >= R + (R & 1)
        double* dst = scol + (int64_t)buf * FM_KP * CP;
        for (int e = threadIdx.x; e < t * (FM_TR / 2); e += 256) {
            const int u = e / (FM_TR / 2), pc = e - u * (FM_TR / 2);
            if (pc < pieces) cp_async16(dst + u * CP + 2 * pc, B.colP + (int64_t)u * B.Rpad + i0 + 2 * pc);
        }
        cp_async_commit();
    };

    // walk of this warp over its 32-row groups: (strip, row block, group) of the group whose loads are issued next
    int64_t n_strip = it0 / n_rb, n_rbk = it0 - n_strip * n_rb, n_it = it0;
    int n_g = 0;
    int64_t skip_strip = -1;
    unsigned skip = 0;  // bit ni: this thread's column pair of tile ni is a pivot column pair or lies beyond the tableau
    struct Grp {
        double* base;   // element (first row of the group + li, first column of the warp + 2 lk)
        unsigned off;   // bit 4 mi + ni: do not touch this thread's pair of tile (mi, ni)
    };
    auto next_group = [&]() {
        Grp G;
        const int64_t row0 = n_rbk * FM_TR + n_g * 32;
        const int64_t col0 = n_strip * FM_SW + warp * 32;
        unsigned rowoff = 0;  // bit mi: row row0 + 8 mi + li is a pivot row or lies beyond the tableau
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int64_t i = row0 + 8 * mi + li;
            bool bad = i >= R;
            for (int u = 0; u < t; ++u) bad |= (sr[u] == (int32_t)i);
            rowoff |= bad ? (1u << mi) : 0u;
        }
        if (n_strip != skip_strip) {  // a CTA changes strip at most a few times
            skip_strip = n_strip;
            skip = 0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int64_t j = col0 + 8 * ni + 2 * lk;
                bool bad = j >= C;
                for (int u = 0; u < t; ++u) bad |= (ss[u] >= 0 && (ss[u] & ~1) == j);
                skip |= bad ? (1u << ni) : 0u;
            }
        }
        unsigned off = 0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                if (n_it >= it1 || ((rowoff >> mi) & 1u) || ((skip >> ni) & 1u)) off |= 1u << (4 * mi + ni);
        G.off = off;
        G.base = T + (row0 + li) * ld + col0 + 2 * lk;
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
    auto load = [&](const Grp& G, double2 (*v)[4]) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                v[mi][ni] = make_double2(0.0, 0.0);
                if (!((G.off >> (4 * mi + ni)) & 1u))
                    v[mi][ni] = ld_stream(reinterpret_cast<const double2*>(G.base + (int64_t)(8 * mi) * ld + 8 * ni));
            }
    };

    double2 vn[4][4];
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
                for (int c = lane; c < FM_SW; c += 32) {
                    const int64_t jj = strip * FM_SW + c;
                    sq[u * QP + c] = jj < B.Cpad ? -B.qP[(int64_t)u * B.Cpad + jj] : 0.0;
                }
            __syncthreads();
        }
        int64_t nstrip = strip, nrb = rb + 1;
        if (nrb == n_rb) {
            nrb = 0;
            ++nstrip;
        }
        if (it + 1 < it1) stage_cols(nrb, buf ^ 1);
        // operands of this thread: step lk of every k-block; rows li + 8 mi of the group, columns li + 8 ni of the warp
        const double* qa = sq + lk * QP + warp * 32 + li;
        const double* ca = scol + (int64_t)buf * FM_KP * CP + lk * CP + li;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            const Grp gc = gn;
            double2 v[4][4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) v[mi][ni] = vn[mi][ni];
            auto kblock = [&](int kb) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = ca[kb * 4 * CP + g * 32 + 8 * mi];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = qa[kb * 4 * QP + 8 * ni];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(v[mi][ni], a[mi], b[ni]);
            };
            // The first k-block consumes every register of the previous batch of loads, i.e. the wait on their scoreboard
            // happens HERE, before the same load instructions are issued again for the next group.
            kblock(0);
            asm volatile("" ::: "memory");
            gn = next_group();
            load(gn, vn);
            asm volatile("" ::: "memory");
#pragma unroll 2
            for (int kb = 1; kb < t4 / 4; ++kb) kblock(kb);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    if (!((gc.off >> (4 * mi + ni)) & 1u))
                        st_stream(reinterpret_cast<double2*>(gc.base + (int64_t)(8 * mi) * ld + 8 * ni), v[mi][ni]);
        }
        strip = nstrip;
        rb = nrb;
    }
}

}  // namespace b200lp
