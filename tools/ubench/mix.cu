// Does a DFMA hold the sub-partition's dispatch port for both of its two pipe cycles?  W warps per SM, each with C independent
// DFMA chains, optionally interleaved with K independent integer IMADs (or FFMAs) per DFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int C, int K, int KIND>
__global__ void mix(double *out, int iters, double a, double b, int ia, int ib) {
    double x[C];
    int y[8];
    float z[8];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x * 1e-3 + c;
#pragma unroll
    for (int k = 0; k < 8; ++k) { y[k] = threadIdx.x + k; z[k] = threadIdx.x * 0.5f + k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                x[c] = fma(x[c], a, b);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (KIND == 0) y[(c * K + k) & 7] = y[(c * K + k) & 7] * ia + ib;
                    else z[(c * K + k) & 7] = fmaf(z[(c * K + k) & 7], (float)a, (float)b);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += y[k] + z[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int C, int K, int KIND>
void run(int warps, const char *name) {
    double *out; cudaMalloc(&out, 148 * 1024 * 8);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<C, K, KIND><<<148, warps * 32>>>(out, 100, 1.0000001, 1e-9, 3, 1);
    cudaEventRecord(e0);
    mix<C, K, KIND><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9, 3, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma_per_smsp = (double)iters * 8 * C * warps / 4.0;  // warp-level DFMAs per sub-partition
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-34s warps/SM %2d: %.2f cycles per DFMA per sub-partition (other instr per DFMA: %d)\n", name, warps, cyc / dfma_per_smsp, K);
    cudaFree(out);
}
int main() {
    for (int w : {4, 16}) {
        if (w == 4) { run<4, 0, 0>(w, "DFMA x4 chains"); run<4, 1, 0>(w, "DFMA x4 + 1 IMAD"); run<4, 2, 0>(w, "DFMA x4 + 2 IMAD"); run<4, 1, 1>(w, "DFMA x4 + 1 FFMA"); run<4, 2, 1>(w, "DFMA x4 + 2 FFMA"); }
        else { run<1, 0, 0>(w, "DFMA x1 chain"); run<1, 1, 0>(w, "DFMA x1 + 1 IMAD"); run<1, 2, 0>(w, "DFMA x1 + 2 IMAD"); run<1, 1, 1>(w, "DFMA x1 + 1 FFMA"); run<1, 2, 1>(w, "DFMA x1 + 2 FFMA"); run<2, 1, 0>(w, "DFMA x2 + 1 IMAD"); run<2, 2, 0>(w, "DFMA x2 + 2 IMAD"); }
    }
}
