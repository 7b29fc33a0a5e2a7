// pingpong.cu -- cost of one owner<->group hand-off pair per RHS evaluation (kernel C's synchronisation skeleton, no work):
// 4 owner warps + 12 group warps per CTA, one CTA per SM.  Variants: named barriers (bar.arrive / bar.sync), mbarrier
// (arrive + try_wait.parity spin), shared-memory counters polled with volatile loads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pingpong tools/ubench/pingpong.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long *out, int iters, int work) {
    __shared__ uint64_t mb[2];
    __shared__ volatile int cnt[2];
    __shared__ double dummy[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool owner = warp < 4;
    if (threadIdx.x == 0) { mb_init(&mb[0], 12); mb_init(&mb[1], 4); cnt[0] = 0; cnt[1] = 0; }
    __syncthreads();
    double acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            if (owner) { bar_arrive(2, 512); bar_sync(1, 512); }
            else { bar_sync(2, 512); for (int i = 0; i < work; ++i) acc = acc * 1.0000001 + 1e-9; dummy[threadIdx.x] = acc; bar_arrive(1, 512); }
        } else if (MODE == 1) {  // mbarrier: one elected lane per warp arrives, all lanes wait
            if (owner) { __syncwarp(); if (lane == 0) mb_arrive(&mb[1]); mb_wait(&mb[0], it & 1); }
            else { mb_wait(&mb[1], it & 1); for (int i = 0; i < work; ++i) acc = acc * 1.0000001 + 1e-9; dummy[threadIdx.x] = acc; __syncwarp(); if (lane == 0) mb_arrive(&mb[0]); }
        } else {  // polled counters
            if (owner) { __syncwarp(); if (lane == 0) atomicAdd((int *)&cnt[1], 1); while (cnt[0] < 12 * (it + 1)) {} }
            else { while (cnt[1] < 4 * (it + 1)) {} for (int i = 0; i < work; ++i) acc = acc * 1.0000001 + 1e-9; dummy[threadIdx.x] = acc; __threadfence_block(); __syncwarp(); if (lane == 0) atomicAdd((int *)&cnt[0], 1); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.678) out[0] = 0;
}
int main() {
    long long *d, h[148];
    cudaMalloc(&d, sizeof h);
    const int iters = 4000;
    for (int work = 0; work <= 64; work += 32)
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, 512>>>(d, iters, work);
                if (mode == 1) k<1><<<148, 512>>>(d, iters, work);
                if (mode == 2) k<2><<<148, 512>>>(d, iters, work);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("work %2d mode %d (%s): %.1f cycles per round trip  (%s)\n", work, mode, mode == 0 ? "named barriers" : mode == 1 ? "mbarrier" : "polled counters",
                   (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
