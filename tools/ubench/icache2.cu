// I-cache probe 2: FFMA bodies (1 issue/cycle possible). mode 0: all 4 warps run the same copy; mode 1: each warp its own copy.
#include <cstdio>
#include <cuda_runtime.h>
#define OP8 v0 = __fmaf_rn(v0, a, b); v1 = __fmaf_rn(v1, a, b); v2 = __fmaf_rn(v2, a, b); v3 = __fmaf_rn(v3, a, b); \
            v4 = __fmaf_rn(v4, a, b); v5 = __fmaf_rn(v5, a, b); v6 = __fmaf_rn(v6, a, b); v7 = __fmaf_rn(v7, a, b);
#define R2(x) x x
#define R4(x) R2(x) R2(x)
#define R8(x) R4(x) R4(x)
#define R16(x) R8(x) R8(x)
#define R32(x) R16(x) R16(x)
#define R64(x) R32(x) R32(x)
#define R128(x) R64(x) R64(x)
#define R256(x) R128(x) R128(x)
#define KERNEL(NAME, BODY)                                                                 \
    __global__ void NAME(float *out, long long *cyc, int n, int mode) {                    \
        float a = out[1], b = out[2], v0 = out[0] + threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7; \
        const int w = mode ? (threadIdx.x >> 5) & 3 : 0;                                   \
        long long t0 = clock64();                                                          \
        if (w == 0) { for (int i = 0; i < n; ++i) { BODY } }                               \
        else if (w == 1) { for (int i = 0; i < n; ++i) { BODY v0 += 1.f; } }               \
        else if (w == 2) { for (int i = 0; i < n; ++i) { BODY v1 += 2.f; } }               \
        else { for (int i = 0; i < n; ++i) { BODY v2 += 3.f; } }                           \
        long long t1 = clock64();                                                          \
        out[8 + blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7; \
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;                         \
    }
KERNEL(k128, R16(OP8))
KERNEL(k256, R32(OP8))
KERNEL(k512, R64(OP8))
KERNEL(k1024, R128(OP8))
KERNEL(k2048, R256(OP8))
int main() {
    float *d; long long *c; cudaMalloc(&d, 4 * (8 + 148 * 1024)); cudaMalloc(&c, 64);
    float h[4] = {1.0f, 0.999f, 1e-3f, 0}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    long long cy;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 8; warps *= 2) {
#define RUN(K, N) { int n = 400000 / N + 1; K<<<148, 32 * warps>>>(d, c, n, mode); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); \
    printf("mode %d warps/SM %2d body %4d instrs (%3d KB): %.2f cycles per instr per warp, per SMSP %.2f\n", mode, warps, N, N * 16 / 1024, (double)cy / ((double)n * N), (double)cy / ((double)n * N) / (warps / 4)); }
            RUN(k128, 128) RUN(k256, 256) RUN(k512, 512) RUN(k1024, 1024) RUN(k2048, 2048)
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
