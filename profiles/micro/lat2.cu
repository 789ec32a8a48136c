// Second round of microbenchmarks: cost of the rank-1 window update as issued by ONE warp
// (broadcast LDS.128 batches + 26 DFMA), LDS broadcast throughput, DMMA m8n8k4 latency/throughput.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ void lds128(double &a, double &b, unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ __align__(16) double sm[2048];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    double v = seed + lane * 1e-9;
    double acc[26];
#pragma unroll
    for (int q = 0; q < 26; ++q) acc[q] = v + q;
    long long t0, t1;
    // A. batched: 13 LDS.128 then 26 DFMA
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double w[26];
        const unsigned a = base + (i & 31) * 256;
#pragma unroll
        for (int q = 0; q < 13; ++q) lds128(w[2 * q], w[2 * q + 1], a + 16 * q);
#pragma unroll
        for (int q = 0; q < 26; ++q) acc[q] = fma(-v, w[q], acc[q]);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // B. only the 13 LDS.128 (throughput), results consumed by one add chain at the end
    double s0 = 0, s1 = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double w[26];
        const unsigned a = base + (i & 31) * 256;
#pragma unroll
        for (int q = 0; q < 13; ++q) lds128(w[2 * q], w[2 * q + 1], a + 16 * q);
        s0 += w[0] + w[25];
        s1 += w[7];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // C. only 26 independent DFMA from registers
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int q = 0; q < 26; ++q) acc[q] = fma(-v, s0, acc[q]);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // D. 26 LDS.64 broadcast batched + 26 DFMA
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double w[26];
        const unsigned a = base + (i & 31) * 256;
#pragma unroll
        for (int q = 0; q < 26; ++q) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w[q]) : "r"(a + 8 * q) : "memory");
#pragma unroll
        for (int q = 0; q < 26; ++q) acc[q] = fma(-v, w[q], acc[q]);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // E. DMMA m8n8k4 dependent chain
    double d0 = v, d1 = v + 1;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) dmma(d0, d1, 1e-3, 1e-3);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // F. DMMA throughput: 10 independent accumulator tiles
    double e[20];
#pragma unroll
    for (int q = 0; q < 20; ++q) e[q] = v + q;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int q = 0; q < 10; ++q) dmma(e[2 * q], e[2 * q + 1], 1e-3, 1e-3);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // G. per-lane (non-broadcast, conflict-free) LDS.64 x 8 batched
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double w[8];
        const unsigned a = base + ((i & 3) * 256 + lane) * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w[q]) : "r"(a + 256 * q) : "memory");
        s0 += (w[0] + w[1]) + (w[2] + w[3]) + (w[4] + w[5]) + (w[6] + w[7]);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // H. 32-bit shuffle dependent chain
    int iv = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) iv = __shfl_sync(0xffffffffu, iv, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    double r = s0 + s1 + d0 + d1 + iv;
#pragma unroll
    for (int q = 0; q < 26; ++q) r += acc[q];
#pragma unroll
    for (int q = 0; q < 20; ++q) r += e[q];
    out[threadIdx.x] = r;
}
int main() {
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 16 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char *nm[] = {"A 13 LDS.128 bcast + 26 DFMA", "B 13 LDS.128 bcast only", "C 26 DFMA only", "D 26 LDS.64 bcast + 26 DFMA",
                        "E DMMA m8n8k4 dependent", "F 10 indep DMMA m8n8k4", "G 8 per-lane LDS.64 + adds", "H dep SHFL.32"};
    for (int i = 0; i < 8; ++i) printf("%-36s %8.2f cycles/iter\n", nm[i], (double)h[i] / N);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
