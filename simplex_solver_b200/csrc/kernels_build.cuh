// kernels_build.cuh -- producing the initial tableau on the device.
//   k_generate : synthetic dense LP of BASELINE configs 4/5 from a counter-based generator (the 137 GB tableau
//                of config 5 cannot be staged through the host), bit-identical to oracle/simplex_oracle.c.
//   k_build_*  : initial condensed tableau of "min c'x, A_i x (op_i) b_i, x >= 0" -- the arrays that the
//                reference hands to linprog (/root/reference/app/controllers/solver_controller.py:122-170).
#pragma once
#include "common.cuh"

namespace b200lp {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t seed, uint64_t i, uint64_t j) {
    const uint64_t h = mix64(mix64(seed ^ (i * 0xD1342543DE82EF95ull)) + j);
    return (double)(h >> 11) * 0x1.0p-53;
}
constexpr uint64_t KEY_RHS = 0xFFFFFFFFull;

// value of the generated LP "max c'x, Ax <= b" at tableau row i and the column owned by `label` (<0: RHS)
__device__ __forceinline__ double gen_entry(uint64_t seed, int64_t m, int64_t n_total, int64_t i, int64_t label) {
    if (i < m) {
        if (label >= 0) return u01(seed, (uint64_t)i, (uint64_t)label);
        const double u = u01(seed, (uint64_t)i, KEY_RHS);
        double t = __dmul_rn(0.9, u);
        t = __dadd_rn(t, 0.1);
        return __dadd_rn(__dmul_rn(0.25, (double)n_total), t);
    }
    if (label >= 0) {
        const double u = u01(seed, (uint64_t)m, (uint64_t)label);
        double t = __dmul_rn(0.9, u);
        t = __dadd_rn(t, 0.1);
        return -t;
    }
    return 0.0;
}

// one thread per column pair, grid.y strides over rows: coalesced 128-bit stores
__global__ void __launch_bounds__(256)
k_generate(double* __restrict__ T, int64_t m, int64_t R, int64_t C, int64_t ld, uint64_t seed, int64_t n_total,
           int64_t lab0, int32_t* __restrict__ rowlab, int32_t* __restrict__ collab) {
    const int64_t j = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (blockIdx.x == 0) {
        for (int64_t i = blockIdx.y * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.y * blockDim.x)
            rowlab[i] = i < m ? (int32_t)(n_total + i) : -1;
    }
    if (j >= ld) return;
    const int64_t labx = (j < C - 1) ? lab0 + j : -1;
    const int64_t laby = (j + 1 < C - 1) ? lab0 + j + 1 : -1;
    for (int64_t i = blockIdx.y; i < R; i += gridDim.y) {
        double2 v;
        v.x = (j < C) ? gen_entry(seed, m, n_total, i, labx) : 0.0;
        v.y = (j + 1 < C) ? gen_entry(seed, m, n_total, i, laby) : 0.0;
        *reinterpret_cast<double2*>(T + i * ld + j) = v;
    }
    if (blockIdx.y == 0) {
        if (j < C) collab[j] = (int32_t)labx;
        if (j + 1 < C) collab[j + 1] = (int32_t)laby;
    }
}

// Per-row build info computed on the host from (ops, sign of b): bit0 = row negated, bits 1-2 = operator
// after normalisation, surplus = position of the row's surplus column or -1.
struct RowInfo {
    int32_t flags;
    int32_t surplus;
};

// constraint rows and the objective row; one thread per column pair
__global__ void __launch_bounds__(256)
k_build_rows(double* __restrict__ T, int64_t m, int64_t n, int64_t C, int64_t ld, const double* __restrict__ A,
             int64_t lda, const double* __restrict__ b, const double* __restrict__ c,
             const RowInfo* __restrict__ info) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    for (int64_t i = blockIdx.y; i <= m; i += gridDim.y) {
        double v = 0.0;
        if (i < m) {
            const RowInfo ri = info[i];
            const bool neg = ri.flags & 1;
            if (j < n) {
                const double a = A[i * lda + j];
                v = neg ? -a : a;
            } else if (j == C - 1) {
                const double bi = b[i];
                v = neg ? -bi : bi;
            } else if (j == ri.surplus) {
                v = -1.0;
            }
        } else {
            if (j < n) v = c[j];
        }
        T[i * ld + j] = v;
    }
}

// phase-1 row: w_j = -(sum over artificial rows, ascending, plain adds) -- same order as the oracle
__global__ void __launch_bounds__(256)
k_build_phase1_row(double* __restrict__ T, int64_t m, int64_t C, int64_t ld, const int32_t* __restrict__ rowlab,
                   int32_t art_base) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    double acc = 0.0;
    if (j < C) {
        for (int64_t i = 0; i < m; ++i)
            if (rowlab[i] >= art_base) acc = __dadd_rn(acc, T[i * ld + j]);
    }
    T[(m + 1) * ld + j] = (j < C) ? -acc : 0.0;
}

__global__ void __launch_bounds__(256)
k_read_solution(const double* __restrict__ T, int64_t m, int64_t C, int64_t ld, const int32_t* __restrict__ rowlab,
                int64_t n_struct, double* __restrict__ x, double* __restrict__ fun) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const int32_t lab = rowlab[i];
        if (lab >= 0 && lab < n_struct) x[lab] = T[i * ld + C - 1];
    }
    if (i == 0) *fun = -T[m * ld + C - 1];
}

}  // namespace b200lp
