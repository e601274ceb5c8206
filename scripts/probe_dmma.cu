// probe_dmma.cu -- is mma.sync.m8n8k4.f64 on sm_100a bit-identical to the sequential chain
//     d = fma(a3, b3, fma(a2, b2, fma(a1, b1, fma(a0, b0, c))))        (k = 0 first)
// that the look-ahead flush applies to every element?  And how fast is it?  Build + run on the GPU box:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/probe_dmma scripts/probe_dmma.cu && /tmp/probe_dmma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// random double with a random exponent in [-span, span] and random sign
__device__ double rnd(uint64_t key, int span) {
    const uint64_t h = mix64(key);
    const double m = 1.0 + (double)(h >> 12) * 0x1.0p-52;
    const int e = (int)((h >> 3) % (uint64_t)(2 * span + 1)) - span;
    return ((h & 1) ? -m : m) * exp2((double)e);
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// counts[0] tiles elements equal to the k = 0..3 chain, [1] to the k = 3..0 chain, [2] to the pairwise tree, [3] total
__global__ void k_check(uint64_t seed, int span, unsigned long long* counts) {
    const int lane = threadIdx.x & 31;
    const uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int gi = lane >> 2, gk = lane & 3;          // A[gi][gk], B[gk][gi'] with gi' = lane >> 2
    const double a = rnd(seed ^ (tile * 131 + 1) * 1000003ull + lane, span);
    const double b = rnd(seed ^ (tile * 131 + 2) * 1000003ull + lane, span);
    const int ci = lane >> 2, cj = (lane & 3) * 2;    // C[ci][cj], C[ci][cj + 1]
    double c0 = rnd(seed ^ (tile * 131 + 3) * 1000003ull + 2 * lane, span);
    double c1 = rnd(seed ^ (tile * 131 + 3) * 1000003ull + 2 * lane + 1, span);
    double d0 = c0, d1 = c1;
    dmma(d0, d1, a, b);
    // reference chains: A[ci][k] lives in lane ci * 4 + k; B[k][cj] in lane cj * 4 + k
    double f0 = c0, f1 = c1, r0 = c0, r1 = c1, p[4], q[4];
    for (int k = 0; k < 4; ++k) {
        const double ak = __shfl_sync(0xffffffffu, a, ci * 4 + k);
        const double b0 = __shfl_sync(0xffffffffu, b, cj * 4 + k), b1 = __shfl_sync(0xffffffffu, b, (cj + 1) * 4 + k);
        f0 = __fma_rn(ak, b0, f0);
        f1 = __fma_rn(ak, b1, f1);
        p[k] = ak * b0;
        q[k] = ak * b1;
    }
    for (int k = 3; k >= 0; --k) {
        const double ak = __shfl_sync(0xffffffffu, a, ci * 4 + k);
        const double b0 = __shfl_sync(0xffffffffu, b, cj * 4 + k), b1 = __shfl_sync(0xffffffffu, b, (cj + 1) * 4 + k);
        r0 = __fma_rn(ak, b0, r0);
        r1 = __fma_rn(ak, b1, r1);
    }
    const double t0 = ((p[0] + p[1]) + (p[2] + p[3])) + c0, t1 = ((q[0] + q[1]) + (q[2] + q[3])) + c1;
    (void)gi; (void)gk;
    unsigned long long e0 = (d0 == f0) + (d1 == f1), e1 = (d0 == r0) + (d1 == r1), e2 = (d0 == t0) + (d1 == t1);
    for (int o = 16; o; o >>= 1) {
        e0 += __shfl_xor_sync(0xffffffffu, e0, o);
        e1 += __shfl_xor_sync(0xffffffffu, e1, o);
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    }
    if (lane == 0) {
        atomicAdd(&counts[0], e0);
        atomicAdd(&counts[1], e1);
        atomicAdd(&counts[2], e2);
        atomicAdd(&counts[3], 64ull);
    }
}

// throughput: 16 independent accumulator tiles per warp, `iters` DMMAs each; and the same flops as DFMA chains
__global__ void k_rate_dmma(int iters, double* out) {
    double d[16][2];
    for (int t = 0; t < 16; ++t) d[t][0] = d[t][1] = threadIdx.x + t;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int t = 0; t < 16; ++t) dmma(d[t][0], d[t][1], a, b);
    double s = 0;
    for (int t = 0; t < 16; ++t) s += d[t][0] + d[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rate_dfma(int iters, double* out) {
    double d[32];
    for (int t = 0; t < 32; ++t) d[t] = threadIdx.x + t;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int t = 0; t < 32; ++t) d[t] = __fma_rn(a, b, d[t]);
    double s = 0;
    for (int t = 0; t < 32; ++t) s += d[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// are the DMMA and DFMA pipes independent?  Half the warps of every CTA run DMMA, the other half DFMA, each as much work
// as in the pure kernels above: if the pipes are separate the mixed kernel takes max(t_dmma, t_dfma) / 2, not the sum / 2
__global__ void k_rate_mixed(int iters, double* out) {
    const bool use_mma = ((threadIdx.x >> 5) & 1) == 0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double s = 0;
    if (use_mma) {
        double d[16][2];
        for (int t = 0; t < 16; ++t) d[t][0] = d[t][1] = threadIdx.x + t;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int t = 0; t < 16; ++t) dmma(d[t][0], d[t][1], a, b);
        for (int t = 0; t < 16; ++t) s += d[t][0] + d[t][1];
    } else {
        double d[32];
        for (int t = 0; t < 32; ++t) d[t] = threadIdx.x + t;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int t = 0; t < 32; ++t) d[t] = __fma_rn(a, b, d[t]);
        for (int t = 0; t < 32; ++t) s += d[t];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    unsigned long long* counts;
    cudaMallocManaged(&counts, 4 * sizeof(unsigned long long));
    for (int span : {0, 1, 8, 30, 60, 300}) {
        for (int i = 0; i < 4; ++i) counts[i] = 0;
        k_check<<<4096, 256>>>(0x1234567ull + span, span, counts);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        printf("exponent span +-%3d: %llu elements; equal to fma chain k=0..3: %llu, k=3..0: %llu, pairwise sum: %llu\n", span,
               counts[3], counts[0], counts[1], counts[2]);
    }
    double* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 4096;
        float ms;
        cudaEventRecord(e0);
        k_rate_dmma<<<148 * 4, 256>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma1 = 148.0 * 4 * 8 * 16 * 256.0 * iters;  // warps x tiles x (8*8*4) FMAs
        printf("DMMA m8n8k4: %.2f T FMA/s (%.3f ms)\n", fma1 / (ms * 1e-3) / 1e12, ms);
        cudaEventRecord(e0);
        k_rate_dfma<<<148 * 4, 256>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma2 = 148.0 * 4 * 256 * 32.0 * iters;
        printf("DFMA       : %.2f T FMA/s (%.3f ms)\n", fma2 / (ms * 1e-3) / 1e12, ms);
        cudaEventRecord(e0);
        k_rate_mixed<<<148 * 4, 256>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mixed (4 warps DMMA + 4 warps DFMA per CTA): %.2f T FMA/s in total (%.3f ms)\n", (fma1 + fma2) / 2 / (ms * 1e-3) / 1e12, ms);
    }
    return 0;
}
