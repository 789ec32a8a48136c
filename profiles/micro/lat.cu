// Latency microbenchmarks for the quantities the warp-synchronous LDL^T design depends on
// (dependent DFMA, SHFL, MUFU.RCP64H, LDS, STS->LDS round trip) on one warp of one SM.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    double a = seed + lane * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 4 independent DFMA chains (throughput, 1 warp)
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    a += a1 + a2 + a3;
    // 3. dependent 64-bit shuffle chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4. dependent rcp.approx.ftz.f64
    a = fabs(a) + 1.5;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5. dependent LDS chain (pointer chasing through doubles holding indices)
    for (int i = lane; i < 1024; i += 32) sm[i] = (double)((i + 33) & 1023);
    __syncwarp();
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. STS -> LDS round trip (store own, read neighbour's)
    double v = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        sm[lane] = v;
        asm volatile("" ::: "memory");
        v = sm[(lane + 1) & 31] + 1.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. dependent DADD, DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = v + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8. dependent FSEL-ish select on double + integer add chain
    int kk = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) kk = (kk - 1) & 31;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 9. shuffle -> DFMA -> shuffle chain (one sweep step)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        double s = __shfl_sync(0xffffffffu, v, i & 31);
        v = fma(-c, s, v);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // 10. 26 independent DFMA + 13 broadcast LDS.128 (one rank-1 update), dependent through v
    double acc[26];
#pragma unroll
    for (int q = 0; q < 26; ++q) acc[q] = v + q;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
        const double2 *c2 = reinterpret_cast<const double2 *>(sm + (i & 15) * 32);
#pragma unroll
        for (int q = 0; q < 13; ++q) {
            double2 w = c2[q];
            acc[2 * q] = fma(-v, w.x, acc[2 * q]);
            acc[2 * q + 1] = fma(-v, w.y, acc[2 * q + 1]);
        }
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[9] = t1 - t0;
    for (int q = 0; q < 26; ++q) v += acc[q];
    out[threadIdx.x] = a + v + idx + kk;
}
int main() {
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 16 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char *nm[] = {"dep DFMA", "4 indep DFMA (per group of 4)", "dep SHFL.64", "dep rcp.approx.f64", "dep LDS.64 (+cvt)",
                        "STS->LDS+DADD round trip", "dep DADD", "dep IADD+LOP", "SHFL+DFMA step", "rank-1 update (26 DFMA + 13 LDS.128)"};
    for (int i = 0; i < 10; ++i) printf("%-40s %8.2f cycles/iter\n", nm[i], (double)h[i] / (i == 9 ? N / 8 : N));
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
