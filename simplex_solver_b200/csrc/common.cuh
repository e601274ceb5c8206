// common.cuh -- device state, reduction keys and small helpers shared by the kernels of libb200lp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lp {

// Loop state, resident in device memory for the whole pivot loop: the host only reads it back every
// `check_every` pivots (no host round trip per pivot).
struct DevState {
    int32_t done;        // loop finished: optimal for this phase, unbounded, or pivot limit
    int32_t status;      // B200LP_STATUS_* (valid when done)
    int32_t have_pivot;  // the current iteration has an entering column (set by price, cleared by ratio)
    int32_t pend;        // row `r` of the last pivot still holds unscaled values (scaled by the next price)
    int32_t s;           // entering column position in this tableau, -1 when it lives in another shard
    int32_t r;           // leaving row
    int32_t enter_lab;   // id of the entering variable
    int32_t leave_lab;   // id of the leaving variable
    int32_t win_rank;    // sharded runs: shard that owns the entering column
    int32_t pad0;
    double p;            // pivot element T[r][s]
    double inv_p;        // 1 / p
    double best_val;     // reduced cost of the entering column (diagnostic)
    long long n_pivots;  // pivots done by this loop
    long long max_pivots;
    unsigned int ticket_price;  // last-block-done counters
    unsigned int ticket_ratio;
    unsigned int ticket_push;   // peer-memory exchange: CTAs of the push kernel that have finished their stores
    unsigned int pad1;
    long long xgen;             // peer-memory exchange: generation of the last completed exchange (never reset)
};

// Candidate of a min-reduction with a total order => the result does not depend on reduction order.
struct Key {
    double v;     // reduced cost (price) or ratio (ratio test)
    int32_t lab;  // variable id used for tie-breaking; INT32_MAX = no candidate
    int32_t pos;  // position (column or row) in the stored tableau
};

#define B200LP_NO_LAB 0x7fffffff

__device__ __forceinline__ Key key_none() {
    Key k;
    k.v = 0.0;
    k.lab = B200LP_NO_LAB;
    k.pos = -1;
    return k;
}

// lexicographic (v, lab): used by Dantzig pricing and by the ratio test
__device__ __forceinline__ bool key_less_val(const Key& a, const Key& b) {
    if (a.lab == B200LP_NO_LAB) return false;
    if (b.lab == B200LP_NO_LAB) return true;
    return a.v < b.v || (a.v == b.v && a.lab < b.lab);
}
// Bland: lowest variable id
__device__ __forceinline__ bool key_less_lab(const Key& a, const Key& b) { return a.lab < b.lab; }

template <bool BY_LABEL>
__device__ __forceinline__ Key key_min(const Key& a, const Key& b) {
    if (BY_LABEL) return key_less_lab(b, a) ? b : a;
    return key_less_val(b, a) ? b : a;
}

// Order-preserving map double -> uint64 (for non-NaN x: a < b  <=>  ord(a) < ord(b)); -0.0 is folded into +0.0 first
// so that the map agrees with operator< on zeros as well.
__device__ __forceinline__ unsigned long long ord_f64(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Lane that holds the warp-wide minimum of the value order (v, variable id), or -1 if no lane has a candidate: three
// 32-bit min-reductions on the REDUX unit (high word and low word of the order-preserving image of v, then the variable
// id -- ids are unique, so exactly one lane is left) instead of five rounds of 4 shuffles + compares.  Measured on B200:
// locating the winner by ballot is cheaper than a fourth reduction over the positions, and a branch that skips the
// second and third reduction when the high words already decide costs more than it saves.
__device__ __forceinline__ int warp_argmin_val_lane(const Key& k) {
    const unsigned FULL = 0xffffffffu;
    const bool has = k.lab != B200LP_NO_LAB;
    const unsigned long long o = ord_f64(k.v);
    const unsigned hi = (unsigned)(o >> 32), lo = (unsigned)o;
    const unsigned mh = __reduce_min_sync(FULL, has ? hi : 0xffffffffu);
    const bool in1 = has && hi == mh;
    const unsigned mlo = __reduce_min_sync(FULL, in1 ? lo : 0xffffffffu);
    const bool in2 = in1 && lo == mlo;
    const unsigned ml = __reduce_min_sync(FULL, in2 ? (unsigned)k.lab : 0xffffffffu);
    return __ffs(__ballot_sync(FULL, in2 && (unsigned)k.lab == ml)) - 1;  // no candidate: empty ballot -> -1
}
// the same for the label order (lowest variable id)
__device__ __forceinline__ int warp_argmin_lab_lane(const Key& k) {
    const unsigned FULL = 0xffffffffu;
    const unsigned ml = __reduce_min_sync(FULL, (unsigned)k.lab);  // NO_LAB is the largest id
    if (ml == (unsigned)B200LP_NO_LAB) return -1;
    return __ffs(__ballot_sync(FULL, (unsigned)k.lab == ml)) - 1;
}

// Warp-wide min of a Key under the same total orders as key_min<>.  Every lane returns the winning Key (key_none() if
// no lane has a candidate).
template <bool BY_LABEL>
__device__ __forceinline__ Key warp_key_min(Key k) {
    const unsigned FULL = 0xffffffffu;
    const int src = BY_LABEL ? warp_argmin_lab_lane(k) : warp_argmin_val_lane(k);
    if (src < 0) return key_none();
    Key r;
    r.v = __shfl_sync(FULL, k.v, src);
    r.lab = __shfl_sync(FULL, k.lab, src);
    r.pos = __shfl_sync(FULL, k.pos, src);
    return r;
}

// First stage of a CTA-wide reduction: the warp's winner is written to *dst by the lane that holds it (no broadcast
// back to the lanes); key_none() if the warp has no candidate.
template <bool BY_LABEL>
__device__ __forceinline__ void warp_key_min_store(const Key& k, Key* dst) {
    const int lane = threadIdx.x & 31;
    const int src = BY_LABEL ? warp_argmin_lab_lane(k) : warp_argmin_val_lane(k);
    if (src < 0) {
        if (lane == 0) *dst = key_none();
    } else if (lane == src) {
        *dst = k;
    }
}

// CTA-wide min of a Key; result valid in every thread of warp 0.  smem: one Key per warp.
template <bool BY_LABEL>
__device__ __forceinline__ Key block_key_min(Key k, Key* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    k = warp_key_min<BY_LABEL>(k);
    if (lane == 0) smem[warp] = k;
    __syncthreads();
    if (warp == 0) {
        k = lane < nwarp ? smem[lane] : key_none();
        k = warp_key_min<BY_LABEL>(k);
    }
    return k;
}

// CTA-wide min of a Key with ONE __syncthreads, result valid in EVERY thread: the per-warp winners are stored by the
// lanes that hold them and every warp reduces them again on its own.  For the latency-bound persistent loops, where
// the two barriers of block_key_min plus a third for a broadcast through shared memory sit on the critical path of a
// pivot.  `slots` (one Key per warp, at most 32) must not be written again before another CTA-wide barrier has passed:
// give every call site of a loop its own array.
template <bool BY_LABEL>
__device__ __forceinline__ Key block_key_min_all(const Key& k, Key* slots) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    warp_key_min_store<BY_LABEL>(k, slots + warp);
    __syncthreads();
    return warp_key_min<BY_LABEL>(lane < nwarp ? slots[lane] : key_none());
}

// streaming (evict-first) 128-bit accesses for tableau elements that are touched once per pivot
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double2* p, double2 v) { __stcs(p, v); }

}  // namespace b200lp
