// Per-group micro-benchmark of kernel B's flux groups: (a) compile and read instruction counts from SASS,
// (b) run each group alone (one warp per CTA, 148 CTAs) and report cycles per evaluation -- the dependency-bound
// floor of a lone warp, to compare with the in-kernel per-phase timings (GLG_PROFILE_GROUPS).
#include <cstdio>
#include <vector>
#include "../../greenlight-gym2_b200/csrc/glg_roles.cuh"
template <int G>
__global__ void k_group(const __grid_constant__ GlgUniform U, double *xs, double *part, double *Hs, int n, long long *cyc) {
    const int lane = threadIdx.x & 31;
    __shared__ double s_xs[28 * 32], s_H[H_COUNT * 32], s_part[80 * 32];
    for (int i = threadIdx.x; i < 28 * 32; i += blockDim.x) s_xs[i] = xs[i];
    for (int i = threadIdx.x; i < H_COUNT * 32; i += blockDim.x) s_H[i] = Hs[i];
    __syncthreads();
    GlgColT<double, 32> Hc{s_H + lane};
    const GlgXsCol<double> X{s_xs + lane};
    part = s_part;
    for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) s_part[i] = 0.0;
    double u[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5};
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        glg_run_group<G, false>(U, GlgConstView{U.C}, Hc, u, X, part + lane);
        __syncthreads();
    }
    long long t1 = clock64();
    double *gpart = Hs + H_COUNT * 32;  // sink so the stores stay live
    for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) gpart[i] = s_part[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / n;
}
template <int G>
void run(const GlgUniform &U, double *xs, double *part, double *Hs, long long *cyc) {
    k_group<G><<<148, 32>>>(U, xs, part, Hs, 2000, cyc);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("group %d alone: %lld cycles per evaluation (%s)\n", G, c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    GlgUniform U;
    double p[GLG_NP];
    FILE *f = fopen("tools/ubench/params.bin", "rb");
    if (!f || fread(p, 8, GLG_NP, f) != GLG_NP) { printf("need tools/ubench/params.bin (208 doubles)\n"); return 1; }
    fclose(f);
    for (int i = 0; i < GLG_NP; ++i) U.P[i] = p[i];
    glg_make_k(p, U.K);
    glg_make_c(p, U.C);
    double x0[28] = {756.86, 756.86, 16.5, 16.5, 20.5, 16.5, 16.5, 16.5, 16.5, 16.5, 16.5, 15.45, 14.4, 13.35, 12.3, 1681.9, 1681.9,
                     16.5, 16.5, 16.5, 16.5, 20.5, 0.0, 9.5283e4, 2.5107e5, 5.5338e4, 3.0978e3, 0.0};
    double uu[6] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5}, d[10] = {128, 10.3, 729.5, 756.86, 3.5, 5.74, 12.3, 0.93, 1, 1};
    double H[H_COUNT];
    glg_hoist(p, uu, d, H);
    std::vector<double> hx(28 * 32), hH(H_COUNT * 32);
    for (int i = 0; i < 28; ++i) for (int l = 0; l < 32; ++l) hx[i * 32 + l] = x0[i] * (1 + 1e-3 * l);
    for (int i = 0; i < H_COUNT; ++i) for (int l = 0; l < 32; ++l) hH[i * 32 + l] = H[i];
    double *xs, *part, *Hs; long long *cyc;
    cudaMalloc(&xs, hx.size() * 8); cudaMalloc(&Hs, (hH.size() + 80 * 32) * 8); cudaMalloc(&part, 80 * 32 * 8); cudaMalloc(&cyc, 8);
    cudaMemcpy(xs, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(Hs, hH.data(), hH.size() * 8, cudaMemcpyHostToDevice);
    run<0>(U, xs, part, Hs, cyc); run<1>(U, xs, part, Hs, cyc); run<2>(U, xs, part, Hs, cyc); run<3>(U, xs, part, Hs, cyc);
    run<4>(U, xs, part, Hs, cyc); run<5>(U, xs, part, Hs, cyc); run<6>(U, xs, part, Hs, cyc); run<7>(U, xs, part, Hs, cyc);
    return 0;
}
