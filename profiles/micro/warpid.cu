// Which hardware warp slots (%warpid) do the warps of two co-resident 128-thread CTAs get?
// Also times a DFMA-bound loop per warp pair to see which slots share a scheduler partition.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 2) k(int *out, int hold) {
    extern __shared__ double sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        out[(blockIdx.x * 4 + w) * 2] = smid;
        out[(blockIdx.x * 4 + w) * 2 + 1] = wid;
    }
    // keep the CTA resident for a while so that both CTAs of an SM coexist
    long long t0 = clock64();
    while (clock64() - t0 < hold) {}
    sm[threadIdx.x] = 0;
}
// two warps run a dependent-free DFMA loop; report cycles -> same partition if ~2x slower
__global__ void __launch_bounds__(128, 1) pair(long long *out, int wa, int wb) {
    int w = threadIdx.x >> 5;
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    __syncthreads();
    long long t0 = clock64();
    if (w == wa || w == wb) {
        for (int i = 0; i < 4096; ++i) {
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
            a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == wa * 32) out[0] = t1 - t0;
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.0) out[1] = 1;
}
int main() {
    int *d; cudaMalloc(&d, 296 * 8 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 105000);
    k<<<296, 128, 105000>>>(d, 2000000);
    int h[296 * 8]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    // print for the first few SMs: blocks and warp ids
    for (int sm = 0; sm < 6; ++sm) {
        printf("SM %d:", sm);
        for (int b = 0; b < 296; ++b) if (h[b * 8] == sm) { printf("  blk %d wid", b); for (int w = 0; w < 4; ++w) printf(" %d", h[(b * 4 + w) * 2 + 1]); }
        printf("\n");
    }
    // histogram: for SMs with two blocks, pattern of wids
    int pat[4] = {0, 0, 0, 0};
    for (int b = 0; b < 296; ++b) { int w0 = h[b * 8 + 1]; pat[(w0 & 3)]++; }
    printf("first-warp slot mod 4 histogram: %d %d %d %d\n", pat[0], pat[1], pat[2], pat[3]);
    int same = 0, tot = 0;
    for (int b = 0; b < 296; ++b) for (int c = b + 1; c < 296; ++c) if (h[b * 8] == h[c * 8]) { tot++; if (((c - b) % 148) == 0) same++; }
    printf("co-resident pairs %d, of which block ids differ by 148: %d\n", tot, same);
    long long *o; cudaMalloc(&o, 16);
    for (int wb = 0; wb < 4; ++wb) { pair<<<1, 128>>>(o, 0, wb); long long r[2]; cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost); printf("warps 0+%d busy: %lld cycles\n", wb, r[0]); }
    return 0;
}
