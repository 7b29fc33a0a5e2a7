// compile-only: one kernel per flux group, to read per-group instruction counts from SASS (balance of kernel B's warps)
#include "../../greenlight-gym2_b200/csrc/glg_roles.cuh"
template <int G>
__global__ void k_group(const __grid_constant__ GlgUniform U, double *xs, double *part, double *Hs, int n) {
    const int lane = threadIdx.x & 31;
    GlgCol<32> Hc{Hs + lane};
    const GlgXsCol X{xs + lane};
    double u[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        glg_run_group<G, false>(U, GlgConstView{U.C}, Hc, u, X, part + lane);
        __syncthreads();
    }
}
template __global__ void k_group<0>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<1>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<2>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<3>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<4>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<5>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<6>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
template __global__ void k_group<7>(const __grid_constant__ GlgUniform, double *, double *, double *, int);
