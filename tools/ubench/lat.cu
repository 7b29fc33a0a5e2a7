// micro-benchmarks: dependent-chain latency of DFMA / MUFU.RCP64H / LDS / BAR.SYNC on sm_100a (1 warp, 1 block)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, long long *cyc, int n) {
    double v = out[0], a = out[1], b = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v = __fma_rn(v, a, b);
    }
    long long t1 = clock64();
    out[3] = v; cyc[0] = t1 - t0;
}
__global__ void k_dfma_ilp(double *out, long long *cyc, int n, int dummy) {
    double v0 = out[0], v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, a = out[1], b = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v0 = __fma_rn(v0, a, b); v1 = __fma_rn(v1, a, b); v2 = __fma_rn(v2, a, b); v3 = __fma_rn(v3, a, b); }
    }
    long long t1 = clock64();
    out[3] = v0 + v1 + v2 + v3; cyc[0] = t1 - t0;
}
__global__ void k_bar(long long *cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double *out, long long *cyc, int n) {
    double v = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); v = r; }
    }
    long long t1 = clock64();
    out[3] = v; cyc[0] = t1 - t0;
}
int main() {
    double *d; long long *c; cudaMalloc(&d, 64); cudaMalloc(&c, 64);
    double h[4] = {1.0, 0.999999, 1e-7, 0}; cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
    long long cy; int n = 1000;
    k_dfma<<<1, 32>>>(d, c, n); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.2f cycles\n", (double)cy / (n * 32));
    for (int w = 1; w <= 8; w *= 2) { k_dfma<<<1, 32 * w * 4>>>(d, c, n); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("  %d warps/SMSP dependent chains: %.2f cycles per DFMA per warp\n", w, (double)cy / (n * 32)); }
    k_dfma_ilp<<<1, 32>>>(d, c, n, 0); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("DFMA ILP4 1 warp: %.2f cycles per DFMA\n", (double)cy / (n * 32));
    k_rcp<<<1, 32>>>(d, c, n); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("MUFU.RCP64H dependent latency: %.2f cycles\n", (double)cy / (n * 16));
    for (int t = 32; t <= 256; t *= 2) { k_bar<<<1, t>>>(c, n); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("__syncthreads %d threads: %.1f cycles\n", t, (double)cy / n); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
