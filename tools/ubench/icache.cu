// I-cache probe: straight-line loop body of N independent-ish DFMAs (ILP 8), 1 or more warps per SMSP, all SMs.
#include <cstdio>
#include <cuda_runtime.h>
#define OP8 v0 = __fma_rn(v0, a, b); v1 = __fma_rn(v1, a, b); v2 = __fma_rn(v2, a, b); v3 = __fma_rn(v3, a, b); \
            v4 = __fma_rn(v4, a, b); v5 = __fma_rn(v5, a, b); v6 = __fma_rn(v6, a, b); v7 = __fma_rn(v7, a, b);
#define R2(x) x x
#define R4(x) R2(x) R2(x)
#define R8(x) R4(x) R4(x)
#define R16(x) R8(x) R8(x)
#define R32(x) R16(x) R16(x)
#define R64(x) R32(x) R32(x)
#define R128(x) R64(x) R64(x)
#define R256(x) R128(x) R128(x)
#define R512(x) R256(x) R256(x)
#define KERNEL(NAME, BODY)                                                                 \
    __global__ void NAME(double *out, long long *cyc, int n) {                             \
        double a = out[1], b = out[2], v0 = out[0] + threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7; \
        long long t0 = clock64();                                                          \
        for (int i = 0; i < n; ++i) { BODY }                                               \
        long long t1 = clock64();                                                          \
        out[8 + blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7; \
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;                         \
    }
KERNEL(k256, R32(OP8))
KERNEL(k512, R64(OP8))
KERNEL(k1024, R128(OP8))
KERNEL(k2048, R256(OP8))
KERNEL(k4096, R512(OP8))
int main() {
    double *d; long long *c; cudaMalloc(&d, 8 * (8 + 148 * 1024)); cudaMalloc(&c, 64);
    double h[4] = {1.0, 0.999999, 1e-7, 0}; cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
    long long cy;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int blocks = 1; blocks <= 148; blocks *= 148) {
#define RUN(K, N) { int n = 200000 / N + 1; K<<<blocks, 32 * warps>>>(d, c, n); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); \
    printf("blocks %3d warps/SM %2d body %4d instrs (%3d KB): %.2f cycles per instr per warp, per SMSP %.2f\n", blocks, warps, N, N * 16 / 1024, (double)cy / ((double)n * N), (double)cy / ((double)n * N) / (warps / 4)); }
            RUN(k256, 256) RUN(k512, 512) RUN(k1024, 1024) RUN(k2048, 2048) RUN(k4096, 4096)
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
