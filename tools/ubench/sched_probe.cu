// Scheduling probe: which code shape makes ptxas give up the interleaved order of independent chains?
// Compile only (cuobjdump + tools/sass_dep_stats.py); variants selected with -DVARIANT=n.
#include "../../greenlight-gym2_b200/csrc/glg_roles.cuh"
#ifndef VARIANT
#define VARIANT 1
#endif
#if VARIANT == 1  // all 8 groups behind a switch on the warp index, nothing else live
__global__ void __launch_bounds__(256, 2) k_probe(const __grid_constant__ GlgUniform U, double *xs, double *Hs, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double s_xs[29 * 32], s_H[H_COUNT * 32], s_part[80 * 32];
    for (int i = threadIdx.x; i < 28 * 32; i += blockDim.x) s_xs[i] = xs[i];
    for (int i = threadIdx.x; i < H_COUNT * 32; i += blockDim.x) s_H[i] = Hs[i];
    __syncthreads();
    GlgColT<double, 32> Hc{s_H + lane};
    const GlgXsCol<double> X{s_xs + lane};
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        glg_run_warp_groups<false, 8>(warp, U, GlgConstView{U.C}, Hc, u, X, s_part + lane);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) Hs[H_COUNT * 32 + i] = s_part[i];
}
#elif VARIANT == 2  // groups + owner phase (the real loop), minimal prologue
__global__ void __launch_bounds__(256, 2) k_probe(const __grid_constant__ GlgUniform U, double *xs, double *Hs, int n, double h) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double s_xs[29 * 32], s_H[H_COUNT * 32], s_part[80 * 32];
    for (int i = threadIdx.x; i < 28 * 32; i += blockDim.x) s_xs[i] = xs[i];
    for (int i = threadIdx.x; i < H_COUNT * 32; i += blockDim.x) s_H[i] = Hs[i];
    __syncthreads();
    GlgColT<double, 32> Hc{s_H + lane};
    const GlgXsCol<double> X{s_xs + lane};
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
    GlgOwnerRegs<8> own;
    glg_owner_setup<8>(U.K, warp, own);
    double xo[4], acc[4];
    for (int j = 0; j < 4; ++j) { xo[j] = s_xs[j * 32 + lane]; acc[j] = 0.0; }
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        glg_run_warp_groups<false, 8>(warp, U, GlgConstView{U.C}, Hc, u, X, s_part + lane);
        __syncthreads();
        glg_owner_update<8>(own, warp, s_xs + lane, s_part + lane, xo, acc, i & 3, h);
        __syncthreads();
    }
    for (int j = 0; j < 4; ++j) Hs[j * 32 + lane] = xo[j];
}
#endif
#if VARIANT >= 3 && VARIANT <= 6
#if VARIANT == 5 || VARIANT == 6
__constant__ GlgUniform g_U;
template <int G>
__device__ __noinline__ void group_call(const double *xs_col, double *part_col, double *h_col) {
    GlgColT<double, 32> Hc{h_col};
    const GlgXsCol<double> X{xs_col};
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
    glg_run_group<G, false>(g_U, GlgConstView{g_U.C}, Hc, u, X, part_col);
}
#endif
__global__ void __launch_bounds__(256, 2) k_probe(const __grid_constant__ GlgUniform U, double *xs, double *Hs, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double s_xs[29 * 32], s_H[H_COUNT * 32], s_part[80 * 32];
    for (int i = threadIdx.x; i < 28 * 32; i += blockDim.x) s_xs[i] = xs[i];
    for (int i = threadIdx.x; i < H_COUNT * 32; i += blockDim.x) s_H[i] = Hs[i];
    __syncthreads();
    GlgColT<double, 32> Hc{s_H + lane};
    const GlgXsCol<double> X{s_xs + lane};
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#if VARIANT == 3
        if (warp == 0) glg_run_group<6, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane);
        else glg_run_group<7, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane);
#elif VARIANT == 4
        glg_run_group<6, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane);
#elif VARIANT == 5
        switch (warp) {
            case 0: group_call<0>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 1: group_call<1>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 2: group_call<2>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 3: group_call<3>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 4: group_call<4>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 5: group_call<5>(s_xs + lane, s_part + lane, s_H + lane); break;
            case 6: group_call<6>(s_xs + lane, s_part + lane, s_H + lane); break;
            default: group_call<7>(s_xs + lane, s_part + lane, s_H + lane); break;
        }
#endif
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) Hs[H_COUNT * 32 + i] = s_part[i];
}
#endif
#if VARIANT >= 10  // VARIANT = 10 + n: the first n groups inlined behind a switch (where does ptxas stop interleaving?)
__global__ void __launch_bounds__(256, 2) k_probe(const __grid_constant__ GlgUniform U, double *xs, double *Hs, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double s_xs[29 * 32], s_H[H_COUNT * 32], s_part[80 * 32];
    for (int i = threadIdx.x; i < 28 * 32; i += blockDim.x) s_xs[i] = xs[i];
    for (int i = threadIdx.x; i < H_COUNT * 32; i += blockDim.x) s_H[i] = Hs[i];
    __syncthreads();
    GlgColT<double, 32> Hc{s_H + lane};
    const GlgXsCol<double> X{s_xs + lane};
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
    constexpr int NG = VARIANT - 10;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        switch (warp) {
            case 0: if (NG > 0) glg_run_group<6, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 1: if (NG > 1) glg_run_group<7, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 2: if (NG > 2) glg_run_group<5, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 3: if (NG > 3) glg_run_group<4, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 4: if (NG > 4) glg_run_group<2, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 5: if (NG > 5) glg_run_group<0, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            case 6: if (NG > 6) glg_run_group<3, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
            default: if (NG > 7) glg_run_group<1, false>(U, GlgConstView{U.C}, Hc, u, X, s_part + lane); break;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) Hs[H_COUNT * 32 + i] = s_part[i];
}
#endif
