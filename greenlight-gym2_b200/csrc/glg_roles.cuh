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
#ifndef GLG_ROLE_MINBLOCKS
#define GLG_ROLE_MINBLOCKS 4
#endif

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
    static constexpr int kColRows = GLG_NX + (GLG_NPART + 2) + H_COUNT + (NOISY ? C_COUNT : 0);  // +zero slot, +canopy scale
    __host__ __device__ static size_t bytes(int Np) {
        return sizeof(double) * ((size_t)kColRows * GLG_ROLE_LANES + (size_t)(Np + 1) * GLG_ND) + 16 +
               sizeof(int) * (5 * GLG_ROLE_LANES + 4);
    }
};

// Owner phase.  ONE copy of the code for all four warps (the SM's instruction cache holds ~32 KB and the four role
// streams already fill it: tools/ubench/icache2.cu): warp w owns states w, w+4, ..., w+24 and finds, per state, the
// (up to three) contribution slots and the capacity-scale constant through small tables in the constant bank.
struct GlgOwnerTable {
    short slot[GLG_NX][3];  // contribution slots to add (GLG_NPART = an always-zero slot)
    short scale_k[GLG_NX];  // index into K of the scale factor, -1: 1.0, -2: per-lane canopy scale from shared memory
};
__host__ __device__ constexpr GlgOwnerTable glg_make_owner_table() {
    GlgOwnerTable t{};
    for (int i = 0; i < GLG_NX; ++i) {
        int n = 0;
        for (int r = 0; r < GLG_NROLES; ++r)
            if ((glg_role_mask(i) >> r) & 1u) t.slot[i][n++] = (short)glg_part_slot(r, i);
        for (; n < 3; ++n) t.slot[i][n] = (short)GLG_NPART;
        t.scale_k[i] = -1;
    }
    t.scale_k[0] = K_INVCAPCO2AIR; t.scale_k[1] = K_INVCAPCO2TOP; t.scale_k[2] = K_INVCAPAIR; t.scale_k[3] = K_INVCAPTOP;
    t.scale_k[4] = -2; t.scale_k[5] = K_INVCAPCOV; t.scale_k[6] = K_INVCAPCOV; t.scale_k[7] = K_INVCAPTHSCR;
    t.scale_k[8] = K_INVCAPFLR; t.scale_k[9] = K_INVCAPPIPE; t.scale_k[17] = K_INVCAPLAMP; t.scale_k[18] = K_INVCAPINTLAMP;
    t.scale_k[19] = K_INVCAPGROPIPE; t.scale_k[20] = K_INVCAPBLSCR;
    return t;
}
__constant__ GlgOwnerTable glg_owner_table = glg_make_owner_table();

__device__ __forceinline__ void glg_owner_update(const double *Kc, int warp, double *xs_col, const double *part_col,
                                                 const double *can_scale_col, double *xo, double *acc, int stage, double h) {
    const double w = (stage == 1 || stage == 2) ? 2.0 : 1.0;
    const double c = (stage == 2) ? h : 0.5 * h;
    const double c6 = h / 6.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int i = warp + 4 * j;
        const double sum = part_col[glg_owner_table.slot[i][0] * GLG_ROLE_LANES] + part_col[glg_owner_table.slot[i][1] * GLG_ROLE_LANES] +
                           part_col[glg_owner_table.slot[i][2] * GLG_ROLE_LANES];
        const int sk = glg_owner_table.scale_k[i];
        const double scale = sk >= 0 ? Kc[sk] : (sk == -1 ? 1.0 : *can_scale_col);
        const double k = scale * sum;
        // stage 0..2: acc = (stage ? acc : 0) + w k ; xs = x + c k      stage 3: x += h/6 (acc + k) ; xs = x
        const double a_new = (stage == 0 ? 0.0 : acc[j]) + w * k;
        const double x_fin = xo[j] + c6 * a_new;  // at stage 3 w = 1: acc + k
        const double xn = (stage == 3) ? x_fin : xo[j] + c * k;
        acc[j] = a_new;
        if (stage == 3) xo[j] = x_fin;
        xs_col[i * GLG_ROLE_LANES] = xn;
    }
}

template <bool GENERAL, bool NOISY>
__global__ void __launch_bounds__(GLG_ROLE_THREADS, GLG_ROLE_MINBLOCKS) glg_step_roles_kernel(const __grid_constant__ GlgUniform U,
                                                                         const __grid_constant__ GlgStepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NL = GLG_ROLE_LANES;
    double *s_wtile = reinterpret_cast<double *>(smem_raw);  // [(Np+1)][10]
    double *s_xs = s_wtile + (size_t)(A.Np + 1) * GLG_ND;     // [28][32]
    double *s_part = s_xs + GLG_NX * NL;                      // [GLG_NPART][32]
    double *s_H = s_part + (GLG_NPART + 2) * NL;              // [H_COUNT][32]; part slot GLG_NPART = 0, GLG_NPART+1 = canopy scale
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
    // a CTA owns A.role_lanes (<= 32) consecutive envs; the remaining lanes are padding (see glg_capi.cu: for small
    // batches fewer envs per CTA means more CTAs, i.e. more warps to hide the FP64 dependency latency)
    const int e = blockIdx.x * A.role_lanes + lane;
    const bool active = lane < A.role_lanes && e < A.B;
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
        s_part[GLG_NPART * NL + lane] = 0.0;
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
        if (warp == 0) {
            GlgPartCol<0> pt{part_col};
            // canopy heat capacity scale 1/(capLeaf*LAI) for state 4, from the stage value of cLeaf (before its owner moves it)
            if (NOISY) {
                glg_role_rad<GENERAL>(Kv, Cc, Hc, Pv, u, X, pt);
                part_col[(GLG_NPART + 1) * NL] = U.K[K_INVCAPLEAF] * glg_rcp(Cc[C_SLA] * X[23]);
            } else {
                glg_role_rad<GENERAL>(Kv, GlgConstView{U.C}, Hc, Pv, u, X, pt);
                part_col[(GLG_NPART + 1) * NL] = U.K[K_INVCAPLEAF] * glg_rcp(U.C[C_SLA] * X[23]);
            }
        } else if (warp == 1) {
            GlgPartCol<1> pt{part_col};
            if (NOISY) glg_role_air(Kv, Cc, Hc, X, pt);
            else glg_role_air(Kv, GlgConstView{U.C}, Hc, X, pt);
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
        glg_owner_update(U.K, warp, xs_col, part_col, part_col + (GLG_NPART + 1) * NL, xo, acc, stage, h);
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
    glg_write_forecast(A, A.role_lanes, s_tbl, s_k, s_tbl_t, s_k_t, s_wtile, uniform, bk, bt);
}
