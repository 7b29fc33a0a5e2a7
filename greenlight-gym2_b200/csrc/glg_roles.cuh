// glg_roles.cuh -- step kernel B: warp-specialised evaluation of the GreenLight RHS.
//
// One CTA = 32 envs x GLG_NROLES warps.  Lane l of every warp works on env (32*blockIdx.x + l); warp r evaluates
// flux group r of the RHS (glg_model.h: RAD / AIR / VAP / CROP) for those 32 envs, so the four groups of one env run
// concurrently on the SM's four sub-partitions.  Why: with one thread per env a 4096-env batch (BASELINE config 2)
// is 128 warps on 592 SM sub-partitions and each lone warp is bound by FP64 dependency latency (ncu: "wait" stall
// dominant, FP64 pipe 17 % busy); splitting the RHS quadruples the resident warps for the same batch, and because a
// role needs a quarter of the hoisted constants and owns a quarter of the RK4 state, its register and
// shared-memory footprint drops enough for 16 resident warps per SM on large batches.
//
// Data flow per RHS evaluation (4 * n_sub per env-step), all through shared memory, two CTA barriers:
//   xs[28][32]        stage state, written by the owner of each state
//   part[slot][32]    role r's contribution to state i (slot = glg_part_slot(r, i)); 45 (role,state) pairs
//   role phase   : every warp reads the xs it needs, evaluates its flux group, stores its contributions
//   -- barrier --
//   owner phase  : state i is owned by warp (i & 3): k_i = scale_i * sum_r part[r][i]; RK4 stage update; xs[i] <- new
//   -- barrier --
// The env-level work (S1 control update, S2 noise, hoisting, S3-S8 epilogue) is done by warp 0 (lane = env) with the
// same device functions as kernel A, so both kernels share one definition of the step semantics.
#pragma once
#include "glg_kernels.cuh"

#define GLG_ROLE_LANES 32
#define GLG_ROLE_THREADS (GLG_ROLE_LANES * GLG_NROLES)

// number of (role, state) pairs before (r, i) in role-major order = slot index of role r's contribution to state i
__host__ __device__ constexpr int glg_part_slot(int r, int i) {
    int n = 0;
    for (int rr = 0; rr < r; ++rr)
        for (int j = 0; j < GLG_NX; ++j) n += (int)((glg_role_mask(j) >> rr) & 1u);
    for (int j = 0; j < i; ++j) n += (int)((glg_role_mask(j) >> r) & 1u);
    return n;
}
constexpr int GLG_NPART = glg_part_slot(GLG_NROLES, 0);

struct GlgXsCol {  // stage-state column of this lane
    const double *b;
    __device__ __forceinline__ double operator[](int i) const { return b[i * GLG_ROLE_LANES]; }
};
template <int R>
struct GlgPartCol {  // contribution slots of role R for this lane
    double *b;
    struct Ref {
        double *p;
        __device__ __forceinline__ void operator=(double v) { *p = v; }
    };
    __device__ __forceinline__ Ref operator[](int i) { return Ref{b + glg_part_slot(R, i) * GLG_ROLE_LANES}; }
};

template <bool NOISY>
struct GlgRoleSmem {
    static constexpr int kColRows = GLG_NX + GLG_NPART + H_COUNT + (NOISY ? C_COUNT : 0);
    __host__ __device__ static size_t bytes(int Np) {
        return sizeof(double) * ((size_t)kColRows * GLG_ROLE_LANES + (size_t)(Np + 1) * GLG_ND) + 16 +
               sizeof(int) * (5 * GLG_ROLE_LANES + 4);
    }
};

// owner phase for warp W: states W, W+4, ..., W+24
template <int W, class KV, class CV>
__device__ __forceinline__ void glg_owner_update(const KV &K, const CV &C, double *xs_col, const double *part_col,
                                                 double *xo, double *acc, int stage, double h, double can_scale) {
    const double w = (stage == 1 || stage == 2) ? 2.0 : 1.0;
    const double c = (stage == 2) ? h : 0.5 * h;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        constexpr int dummy = 0;
        (void)dummy;
        const int i = W + 4 * j;  // compile-time after unrolling
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < GLG_NROLES; ++r)
            if ((glg_role_mask(i) >> r) & 1u) sum += part_col[glg_part_slot(r, i) * GLG_ROLE_LANES];
        const double scale = (i == 4) ? can_scale : glg_state_scale(i, K, C, 0.0);
        const double k = scale * sum;
        double xn;
        if (stage == 3) {
            xn = xo[j] + (h / 6.0) * (acc[j] + k);
            xo[j] = xn;
        } else {
            acc[j] = (stage == 0) ? k : acc[j] + w * k;
            xn = xo[j] + c * k;
        }
        xs_col[i * GLG_ROLE_LANES] = xn;
    }
}

template <bool GENERAL, bool NOISY>
__global__ void __launch_bounds__(GLG_ROLE_THREADS) glg_step_roles_kernel(const __grid_constant__ GlgUniform U,
                                                                         const __grid_constant__ GlgStepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NL = GLG_ROLE_LANES;
    double *s_wtile = reinterpret_cast<double *>(smem_raw);  // [(Np+1)][10]
    double *s_xs = s_wtile + (size_t)(A.Np + 1) * GLG_ND;     // [28][32]
    double *s_part = s_xs + GLG_NX * NL;                      // [GLG_NPART][32]
    double *s_H = s_part + GLG_NPART * NL;                    // [H_COUNT][32]
    double *s_C = s_H + H_COUNT * NL;                         // [C_COUNT][32] (NOISY)
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_xs + (size_t)GlgRoleSmem<NOISY>::kColRows * NL);
    int *s_tbl = reinterpret_cast<int *>(s_bar + 2);
    int *s_k = s_tbl + NL;
    int *s_tbl_t = s_k + NL;
    int *s_k_t = s_tbl_t + NL;
    int *s_bad = s_k_t + NL;
    int *s_misc = s_bad + NL;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x * NL + lane;
    const bool active = e < A.B;
    int k = 0, tbl = 0;
    if (active) {
        k = A.timestep[e];
        tbl = A.table[e];
    }
    const int kw = min(k, A.rows - A.Np - 1);
    int bk, bt;
    const int uniform = glg_stage_weather(A, s_wtile, s_bar, s_misc, active, kw, tbl, bk, bt);
    const double *wrow = uniform ? s_wtile : (A.weather + ((size_t)tbl * A.rows + (size_t)kw) * GLG_ND);

    // ---- prologue: warp 0, lane = env
    double u[GLG_NU], d[GLG_ND];
    double fruit_prev = 0.0;
    unsigned int ctr = 0;
    GlgCol<NL> Hc{s_H + lane};
    GlgCol<NL> Cc{s_C + lane};
    if (warp == 0) {
        double x[GLG_NX];
        if (active) {
            ctr = A.step_ctr[e];
            glg_env_prologue<NOISY>(U, A, e, wrow, ctr, Hc, Cc, x, u, d);
            fruit_prev = x[25];
        } else {
            // padding lanes of the last CTA integrate a copy of a benign state so every warp runs the same loop
            double d0[GLG_ND];
#pragma unroll
            for (int i = 0; i < 7; ++i) d0[i] = A.weather[i];
            glg_init_state(d0, x);
#pragma unroll
            for (int i = 0; i < GLG_NU; ++i) u[i] = 0.0;
            if (NOISY) {
#pragma unroll
                for (int i = 0; i < C_COUNT; ++i) Cc[i] = U.C[i];
            }
            glg_hoist(GlgConstView{U.P}, u, d0, Hc);
        }
#pragma unroll
        for (int i = 0; i < GLG_NX; ++i) s_xs[i * NL + lane] = x[i];
        s_bad[lane] = 0;
    }
    if (GENERAL && warp != 0) {
        // the GENERAL extras of role RAD read raw controls; only warp 0 evaluates that role, nothing to do here
    }
    __syncthreads();

    // ---- integration: role phase / owner phase
    double *xs_col = s_xs + lane;
    double *part_col = s_part + lane;
    const GlgXsCol X{xs_col};
    const GlgConstView Kv{U.K};
    const GlgConstView Pv{U.P};
    double xo[7], acc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        xo[j] = xs_col[(warp + 4 * j) * NL];
        acc[j] = 0.0;
    }
    const double h = A.dt / (double)A.n_sub;
    const int n_eval = 4 * A.n_sub;
#pragma unroll 1
    for (int ev = 0; ev < n_eval; ++ev) {
        const int stage = ev & 3;
        double can_scale = 0.0;
        if (warp == 0) {
            GlgPartCol<0> pt{part_col};
            if (NOISY) {
                glg_role_rad<GENERAL>(Kv, Cc, Hc, Pv, u, X, pt);
                can_scale = U.K[K_INVCAPLEAF] * glg_rcp(Cc[C_SLA] * X[23]);
            } else {
                glg_role_rad<GENERAL>(Kv, GlgConstView{U.C}, Hc, Pv, u, X, pt);
                can_scale = U.K[K_INVCAPLEAF] * glg_rcp(U.C[C_SLA] * X[23]);
            }
        } else if (warp == 1) {
            GlgPartCol<1> pt{part_col};
            glg_role_air(Kv, Hc, X, pt);
        } else if (warp == 2) {
            GlgPartCol<2> pt{part_col};
            if (NOISY) glg_role_vap(Kv, Cc, Hc, X, pt);
            else glg_role_vap(Kv, GlgConstView{U.C}, Hc, X, pt);
        } else {
            GlgPartCol<3> pt{part_col};
            if (NOISY) glg_role_crop<GENERAL>(Kv, Cc, Hc, X, pt);
            else glg_role_crop<GENERAL>(Kv, GlgConstView{U.C}, Hc, X, pt);
        }
        __syncthreads();
        if (warp == 0) glg_owner_update<0>(Kv, GlgConstView{U.C}, xs_col, part_col, xo, acc, stage, h, can_scale);
        else if (warp == 1) glg_owner_update<1>(Kv, GlgConstView{U.C}, xs_col, part_col, xo, acc, stage, h, 0.0);
        else if (warp == 2) glg_owner_update<2>(Kv, GlgConstView{U.C}, xs_col, part_col, xo, acc, stage, h, 0.0);
        else glg_owner_update<3>(Kv, GlgConstView{U.C}, xs_col, part_col, xo, acc, stage, h, 0.0);
        __syncthreads();
    }
    {
        int bad = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) bad |= !(fabs(xo[j]) <= 1.79769313486231570e308);
        if (bad) s_bad[lane] = 1;  // benign race: every writer stores 1
    }
    __syncthreads();

    // ---- epilogue: warp 0, lane = env
    if (warp == 0) {
        GlgEnvOut o;
        o.done = 0; o.k_obs = -1; o.tbl_obs = 0; o.k_term = -1; o.tbl_term = 0; o.fin_ret = 0.0; o.fin_len = 0.0;
#pragma unroll
        for (int j = 0; j < GLG_NINFO; ++j) o.fin_info[j] = 0.0;
        const int bad = s_bad[lane];
        if (active) {
            double x[GLG_NX];
#pragma unroll
            for (int i = 0; i < GLG_NX; ++i) x[i] = xs_col[i * NL];
            glg_env_epilogue(U, A, e, k, kw, tbl, wrow, x, fruit_prev, bad, ctr, o);
        }
        s_tbl[lane] = o.tbl_obs;
        s_k[lane] = o.k_obs;
        s_tbl_t[lane] = o.tbl_term;
        s_k_t[lane] = o.k_term;
        glg_stats_reduce(A, active, bad, o);
    }
    __syncthreads();
    glg_write_forecast(A, NL, s_tbl, s_k, s_tbl_t, s_k_t, s_wtile, uniform, bk, bt);
}
