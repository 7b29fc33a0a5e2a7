// glg_roles.cuh -- step kernel B: warp-specialised evaluation of the GreenLight RHS.
//
// One CTA = 32 envs x NR warps (NR = 4 or 8).  Lane l of every warp works on env (32*blockIdx.x + l); each warp
// evaluates 8/NR of the eight flux groups of the RHS (glg_model.h: rad / fir / airflow / conv / screens / cover / photo /
// flows) for those 32 envs, so the groups of one env run concurrently on the SM's sub-partitions.  Why: with one thread
// per env a 4096-env batch (BASELINE config 2) is 128 warps on 592 SM sub-partitions and each lone warp is bound by the
// 8-cycle FP64 dependency latency (ncu: "wait" stall dominant, FP64 pipe 17 % busy); splitting the RHS multiplies the
// resident warps for the same batch, and because a warp needs only its groups' hoisted constants and owns 1/NR of the RK4
// state, the kernel fits 128 registers / 37 KB shared memory per CTA => 16 resident warps per SM on large batches.
//
// Data flow per RHS evaluation (4 * n_sub per env-step), all through shared memory, two CTA barriers:
//   xs[28][32]        stage state, written by the owner of each state
//   part[slot][32]    group g's contribution to state i (slot = glg_part_slot(g, i)); 55 (group,state) pairs
//   group phase  : every warp reads the xs it needs, evaluates its flux groups, stores their contributions
//   -- barrier --
//   owner phase  : state i is owned by warp (i % NR): k_i = scale_i * sum_g part[g][i]; RK4 stage update; xs[i] <- new
//   -- barrier --
// The owner phase is ONE copy of code for all warps (tables in the constant bank): the SM's instruction cache holds
// ~32 KB (tools/ubench/icache2.cu: four warps on 4 x 16 KB of distinct code run at 8 cycles/instruction instead of 1),
// and the group streams already take ~29 KB of it.
// The env-level work (S1 control update, S2 noise, hoisting, S3-S8 epilogue) is done by warp 0 (lane = env) with the
// same device functions as kernel A, so both kernels share one definition of the step semantics.
#pragma once
#include "glg_kernels.cuh"

#define GLG_ROLE_LANES 32

// number of (group, state) pairs before (g, i) in group-major order = slot index of group g's contribution to state i
__host__ __device__ constexpr int glg_part_slot(int g, int i) {
    int n = 0;
    for (int gg = 0; gg < g; ++gg)
        for (int j = 0; j < GLG_NX; ++j) n += (int)((glg_group_mask(j) >> gg) & 1u);
    for (int j = 0; j < i; ++j) n += (int)((glg_group_mask(j) >> g) & 1u);
    return n;
}
constexpr int GLG_NPART = glg_part_slot(GLG_NGROUPS, 0);
constexpr int GLG_SLOT_ZERO = GLG_NPART;          // always 0.0
constexpr int GLG_SLOT_CANSCALE = GLG_NPART + 1;  // canopy capacity scale of the current stage (written by G0's warp)
constexpr int GLG_SLOT_LAMBDA = GLG_NPART + 2;    // harvest rate constant of the current stage state (written by G7's warp)
constexpr int GLG_SLOT_STIFF = GLG_NPART + 3;     // transient-stiffness estimate of the current stage state (written by G2's warp)
constexpr int GLG_NSLOTS = GLG_NPART + 4;
constexpr int GLG_MAXCONTRIB = 4;

template <class T>
struct GlgXsCol {  // stage-state column of this lane
    const T *b;
    __device__ __forceinline__ T operator[](int i) const { return b[i * GLG_ROLE_LANES]; }
};
template <int G, class T>
struct GlgPartCol {  // contribution slots of group G for this lane
    T *b;
    struct Ref {
        T *p;
        __device__ __forceinline__ void operator=(T v) { *p = v; }
    };
    __device__ __forceinline__ Ref operator[](int i) { return Ref{b + glg_part_slot(G, i) * GLG_ROLE_LANES}; }
};

// Owner plan.  States are sorted by their number of contributing groups (descending) and dealt round-robin to the NR
// owner warps: row j of warp w is state order[j * NR + w].  All warps run ONE copy of straight-line owner code, so row j
// loads as many slots as its hungriest state needs (the row's first entry): 4+2+1+1 = 8 shared-memory loads per warp
// with 8 warps (4+3+2+2+1+1+1 = 14 with 4) instead of 4 per state -- the owner phase is bound by shared-memory
// bandwidth (128 B/clk: one 32-lane fp64 load = 2 cycles), and padding loads of the zero slot were half of it.
constexpr int GLG_PLAN_ROWS = 32;
struct GlgOwnerTable {
    short slot[GLG_NX][GLG_MAXCONTRIB];  // contribution slots to add (GLG_SLOT_ZERO pads)
    short scale_k[GLG_NX];               // glg_state_scale_index
    short order[GLG_PLAN_ROWS];          // states by contribution count, descending; -1 pads
    short count[GLG_PLAN_ROWS];
};
__host__ __device__ constexpr int glg_popcount8(unsigned m) {
    int n = 0;
    for (int g = 0; g < GLG_NGROUPS; ++g) n += (int)((m >> g) & 1u);
    return n;
}
__host__ __device__ constexpr GlgOwnerTable glg_make_owner_table() {
    GlgOwnerTable t{};
    for (int i = 0; i < GLG_NX; ++i) {
        int n = 0;
        for (int g = 0; g < GLG_NGROUPS; ++g)
            if ((glg_group_mask(i) >> g) & 1u) t.slot[i][n++] = (short)glg_part_slot(g, i);
        for (; n < GLG_MAXCONTRIB; ++n) t.slot[i][n] = (short)GLG_SLOT_ZERO;
        t.scale_k[i] = (short)glg_state_scale_index(i);
    }
    int n = 0;
    for (int c = GLG_MAXCONTRIB; c >= 1; --c)
        for (int i = 0; i < GLG_NX; ++i)
            if (glg_popcount8(glg_group_mask(i)) == c) {
                t.order[n] = (short)i;
                t.count[n] = (short)c;
                ++n;
            }
    for (; n < GLG_PLAN_ROWS; ++n) {
        t.order[n] = -1;
        t.count[n] = 1;
    }
    return t;
}
__constant__ GlgOwnerTable glg_owner_table = glg_make_owner_table();
template <int NR, int J>
struct GlgRowSlots {  // slots row J loads (compile-time)
    static constexpr int value = glg_make_owner_table().count[J * NR];
};
template <int NR>
struct GlgCanopyPos {  // (warp, row) of the canopy state 4, whose capacity scale changes with the stage LAI
    static constexpr int find() {
        const GlgOwnerTable t = glg_make_owner_table();
        for (int n = 0; n < GLG_PLAN_ROWS; ++n)
            if (t.order[n] == 4) return n;
        return -1;
    }
    static constexpr int warp = find() % NR, row = find() / NR;
};
template <int I>
struct GlgInt {
    static constexpr int value = I;
};
template <int I, int N, class F>
__device__ __forceinline__ void glg_static_for(F &&f) {
    if constexpr (I < N) {
        f(GlgInt<I>{});
        glg_static_for<I + 1, N>(f);
    }
}

#ifndef GLG_NOINLINE_MASK
#define GLG_NOINLINE_MASK 0x00  // see glg_dispatch_group
#endif
// per-warp copy of the owner tables in shared memory (used by the call build, where registers do not survive the calls)
constexpr int GLG_OWNER_TAB_WORDS = 52;
constexpr int GLG_OWNER_TAB_BYTES = GLG_NOINLINE_MASK ? 8 * GLG_OWNER_TAB_WORDS * 4 : 0;
template <class T, bool NOISY>
struct GlgRoleSmem {
    static constexpr int kColRows = (GLG_NX + 1) + GLG_NSLOTS + H_COUNT + (NOISY ? C_COUNT : 0);  // +1: dummy state row
    // weather tile (f64) | final state (f64 [28][32]) | T columns | mbarrier | ints
    __host__ __device__ static size_t col_bytes() { return (sizeof(T) * (size_t)kColRows * GLG_ROLE_LANES + 15) / 16 * 16; }
    __host__ __device__ static size_t bytes(int Np) {
        return sizeof(double) * ((size_t)(Np + 1) * GLG_ND + (size_t)GLG_NX * GLG_ROLE_LANES) + col_bytes() + 16 +
               sizeof(int) * (5 * GLG_ROLE_LANES + 4) + GLG_OWNER_TAB_BYTES;
    }
};

// owner phase: the per-state table entries (shared-memory offsets of the contribution slots, capacity scale) are loop
// invariants: they are fetched from the constant-bank table ONCE into registers (GlgOwnerRegs) -- dynamic constant-bank
// indexing inside the loop cost ~600 cycles per evaluation.  The update itself is branch-free straight-line code (rows
// without a state update a dummy slot), so the NJ load -> add -> scale -> RK4 chains of a warp overlap.
constexpr int GLG_XS_ROWS = GLG_NX + 1;  // row GLG_NX is the dummy state slot
template <int NR>
struct GlgOwnerRegs {
    static constexpr int NJ = (GLG_NX + NR - 1) / NR;
    int off[NJ][GLG_MAXCONTRIB];  // element offsets of the contribution slots in this lane's column (first GlgRowSlots used)
    double scale[NJ];             // capacity scale
    int xs_off[NJ];               // element offset of the state in the xs column (dummy row for padding)
};
template <int NR>
__device__ __forceinline__ void glg_owner_setup(const double *Kc, int warp, GlgOwnerRegs<NR> &o) {
    glg_static_for<0, GlgOwnerRegs<NR>::NJ>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const int i = glg_owner_table.order[j * NR + warp];
        const bool valid = i >= 0;
        const int ii = valid ? i : 0;
#pragma unroll
        for (int c = 0; c < GlgRowSlots<NR, j>::value; ++c)
            o.off[j][c] = (valid ? glg_owner_table.slot[ii][c] : GLG_SLOT_ZERO) * GLG_ROLE_LANES;
        const int sk = glg_owner_table.scale_k[ii];
        o.scale[j] = (valid && sk >= 0) ? Kc[sk] : 1.0;
        o.xs_off[j] = (valid ? i : GLG_NX) * GLG_ROLE_LANES;
    });
}
// shared-memory image of GlgOwnerRegs: words [0,NJ*4) slot offsets, [NJ*4, NJ*5) xs offsets, then NJ doubles (8-byte aligned)
template <int NR>
__device__ __forceinline__ void glg_owner_store(const GlgOwnerRegs<NR> &o, int *tab) {
    constexpr int NJ = GlgOwnerRegs<NR>::NJ;
    glg_static_for<0, NJ>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
#pragma unroll
        for (int c = 0; c < GlgRowSlots<NR, j>::value; ++c) tab[j * 4 + c] = o.off[j][c];
        tab[NJ * 4 + j] = o.xs_off[j];
        reinterpret_cast<double *>(tab + 36)[j] = o.scale[j];
    });
}
template <int NR>
__device__ __forceinline__ void glg_owner_fetch(GlgOwnerRegs<NR> &o, const int *tab) {
    constexpr int NJ = GlgOwnerRegs<NR>::NJ;
    glg_static_for<0, NJ>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
#pragma unroll
        for (int c = 0; c < GlgRowSlots<NR, j>::value; ++c) o.off[j][c] = tab[j * 4 + c];
        o.xs_off[j] = tab[NJ * 4 + j];
        o.scale[j] = reinterpret_cast<const double *>(tab + 36)[j];
    });
}
template <int NR, class T>
__device__ __forceinline__ void glg_owner_update(const GlgOwnerRegs<NR> &o, int warp, T *xs_col, const T *part_col,
                                                 double *xo, double *acc, int stage, double h, double h_sixth) {
    // stage 0..2: acc = (stage ? acc : 0) + w k ; xs = x + c k        stage 3: x += h/6 (acc + k) ; xs = x
    // h_sixth = h / 6.0 is passed in: the division is done once per (micro-)step size, not once per evaluation
    const bool last = stage == 3;
    const double w = (stage == 1 || stage == 2) ? 2.0 : 1.0;
    const double keep = stage == 0 ? 0.0 : 1.0;
    const double m = last ? h_sixth : (stage == 2 ? h : 0.5 * h);
    const double can_scale = (double)part_col[GLG_SLOT_CANSCALE * GLG_ROLE_LANES];
    double sum[GlgOwnerRegs<NR>::NJ];
    glg_static_for<0, GlgOwnerRegs<NR>::NJ>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int n = GlgRowSlots<NR, j>::value;
        const double a = (double)part_col[o.off[j][0]];
        if constexpr (n == 1) sum[j] = a;
        else if constexpr (n == 2) sum[j] = a + (double)part_col[o.off[j][1]];
        else if constexpr (n == 3) sum[j] = (a + (double)part_col[o.off[j][1]]) + (double)part_col[o.off[j][2]];
        else sum[j] = (a + (double)part_col[o.off[j][1]]) + ((double)part_col[o.off[j][2]] + (double)part_col[o.off[j][3]]);
    });
#pragma unroll
    for (int j = 0; j < GlgOwnerRegs<NR>::NJ; ++j) {
        const double sc = (j == GlgCanopyPos<NR>::row && warp == GlgCanopyPos<NR>::warp) ? can_scale : o.scale[j];
        const double k = sc * sum[j];
        const double a_new = glg_fma(w, k, keep * acc[j]);
        const double xn = glg_fma(m, last ? a_new : k, xo[j]);
        acc[j] = a_new;
        xo[j] = last ? xn : xo[j];
        xs_col[o.xs_off[j]] = (T)xn;
    }
}

// evaluates flux group G for this lane
template <class T>
struct GlgKView;
template <>
struct GlgKView<double> {
    typedef GlgConstView type;
    __device__ __forceinline__ static type k(const GlgUniform &U) { return type{U.K}; }
    __device__ __forceinline__ static type c(const GlgUniform &U) { return type{U.C}; }
};
template <>
struct GlgKView<float> {
    typedef GlgConstViewF type;
    __device__ __forceinline__ static type k(const GlgUniform &U) { return type{U.Kf}; }
    __device__ __forceinline__ static type c(const GlgUniform &U) { return type{U.Cf}; }
};

#if defined(GLG_PROFILE_GROUPS) || defined(GLG_PROFILE_MASK)
__device__ int glg_prof_mask_dev = 0x1FF;  // bits 0..7: run group g ; bit 8: run the owner phase (timing experiments only)
#endif
template <int G, bool GENERAL, bool GUARD = true, class T, class CV, class HV>
__device__ __forceinline__ void glg_run_group(const GlgUniform &U, const CV &Cv, const HV &Hc, const double *u, const GlgXsCol<T> &X,
                                              T *part_col) {
#if defined(GLG_PROFILE_GROUPS) || defined(GLG_PROFILE_MASK)
    if (!((glg_prof_mask_dev >> G) & 1)) return;
#endif
    const typename GlgKView<T>::type Kv = GlgKView<T>::k(U);
    const GlgConstView Pv{U.P};
    GlgPartCol<G, T> pt{part_col};
    if (G == 0) part_col[GLG_SLOT_CANSCALE * GLG_ROLE_LANES] = glg_grp_rad<GENERAL>(Kv, Cv, Hc, X, pt);
    else if (G == 1) glg_grp_fir<GENERAL>(Kv, Cv, Hc, Pv, u, X, pt);
    else if (G == 2) {  // the stiffness estimate is only consumed by the guarded loop
        const T lam = glg_grp_airflow(Kv, Hc, X, pt);
        if (GUARD) part_col[GLG_SLOT_STIFF * GLG_ROLE_LANES] = lam;
    }
    else if (G == 3) glg_grp_conv<GENERAL>(Kv, Cv, Hc, Pv, X, pt);
    else if (G == 4) glg_grp_screens(Kv, Hc, X, pt);
    else if (G == 5) glg_grp_cover(Kv, Cv, Hc, X, pt);
    else if (G == 6) glg_grp_photo<GENERAL>(Kv, Cv, Hc, X, pt);
    else part_col[GLG_SLOT_LAMBDA * GLG_ROLE_LANES] = glg_grp_flows(Kv, Cv, X, pt);
}

// Each flux group is a separate (__noinline__) device function.  Inlined into the kernel, ptxas serialises the hand-
// interleaved chains of the math routines once the function holds all eight groups (FP64 producer distance <= 2 for 40 %
// of the instructions vs 10 % when a group is compiled on its own -- tools/ubench/sched_probe.cu, tools/sass_dep_all.py),
// which doubled the latency of the long groups.  A callee cannot see the kernel's __grid_constant__ parameter, so kernel
// B's group functions read the constants from this __constant__ copy (direct c[3][imm] operands, like the kernel
// parameter was).  There is ONE copy per device: the host uploads a handle's table before launching when another handle
// used it last (glg_capi.cu: bind_uniform), after a device synchronise -- alternating handles on one device serialises.
__constant__ GlgUniform glg_uni_c;

template <int G, bool GENERAL, bool NOISY, bool GUARD, class T>
__device__ __noinline__ void glg_group_call(const T *xs_col, T *part_col, T *h_col, T *c_col, double thScr, double blScr) {
    const GlgUniform &U = glg_uni_c;
    const GlgColT<T, GLG_ROLE_LANES> Hc{h_col};
    const GlgXsCol<T> X{xs_col};
    const double u[GLG_NU] = {0.0, 0.0, thScr, 0.0, 0.0, blScr};  // only G1's GENERAL terms read the raw screen controls
    if (NOISY) {
        const GlgColT<T, GLG_ROLE_LANES> Cc{c_col};
        glg_run_group<G, GENERAL, GUARD>(U, Cc, Hc, u, X, part_col);
    } else {
        glg_run_group<G, GENERAL, GUARD>(U, GlgKView<T>::c(U), Hc, u, X, part_col);
    }
}

// GLG_NOINLINE_MASK: bit g set = group g runs as a __noinline__ call, else inlined into the kernel (default: all
// inlined).  Measured on B200, B = 4096 (profiles/r1_noinline_groups.txt): in-kernel latency of a group running alone,
// cycles -- inlined G0 693, G1 464, G2 723, G3 643, G4 722, G5 895, G6 1058, G7 1010; as calls G0 968, G4 564, G5 600,
// G6 546, G7 680.  The calls keep the interleaved order and halve the long groups' latency, but the step gets SLOWER
// (2.02 ms all inlined, 2.24 ms all calls, 2.39 ms calls for G4..G7 only): two warps share each SM sub-partition's FP64
// pipe (1148 DFMA-class instructions per evaluation = 631 cycles per sub-partition at 2.2 cycles each), and once both
// have ILP they queue on it.  Kept as an experiment switch for the round-2 work on group balance.

template <int G, bool GENERAL, bool NOISY, bool GUARD, class T, class CV>
__device__ __forceinline__ void glg_dispatch_group(const GlgUniform &U, const CV &Cv, const T *xs_col, T *part_col, T *h_col,
                                                   T *c_col, const double *u) {
    if ((GLG_NOINLINE_MASK >> G) & 1) {
        glg_group_call<G, GENERAL, NOISY, GUARD, T>(xs_col, part_col, h_col, c_col, u[2], u[5]);
    } else {
        const GlgColT<T, GLG_ROLE_LANES> Hc{h_col};
        const GlgXsCol<T> X{xs_col};
        glg_run_group<G, GENERAL, GUARD>(U, Cv, Hc, u, X, part_col);
    }
}

template <bool GENERAL, bool NOISY, bool GUARD, int NR, class T, class CV>
__device__ __forceinline__ void glg_run_warp_groups(int warp, const GlgUniform &U, const CV &Cv, const T *xs_col, T *part_col,
                                                    T *h_col, T *c_col, const double *u) {
#define GLG_CALL(G) glg_dispatch_group<G, GENERAL, NOISY, GUARD, T>(U, Cv, xs_col, part_col, h_col, c_col, u)
    if (NR == 8) {
        switch (warp) {
            case 0: GLG_CALL(0); break;
            case 1: GLG_CALL(1); break;
            case 2: GLG_CALL(2); break;
            case 3: GLG_CALL(3); break;
            case 4: GLG_CALL(4); break;
            case 5: GLG_CALL(5); break;
            case 6: GLG_CALL(6); break;
            default: GLG_CALL(7); break;
        }
    } else {  // NR == 4: pairs balanced by measured latency (rad+airflow, fir+conv, screens+cover, photo+flows)
        switch (warp) {
            case 0: GLG_CALL(0); GLG_CALL(2); break;
            case 1: GLG_CALL(1); GLG_CALL(3); break;
            case 2: GLG_CALL(4); GLG_CALL(5); break;
            default: GLG_CALL(6); GLG_CALL(7); break;
        }
    }
#undef GLG_CALL
}

// GRADED: compile the guarded (micro-stepping) loop although the parameters are nominal (NOISY variants always have it).
template <class T, bool GENERAL, bool NOISY, int NR, bool GRADED = false>
__global__ void __launch_bounds__(32 * NR, 16 / NR) glg_step_roles_kernel(const __grid_constant__ GlgUniform U,
                                                                          const __grid_constant__ GlgStepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NL = GLG_ROLE_LANES;
    constexpr int NJ = (GLG_NX + NR - 1) / NR;
    double *s_wtile = reinterpret_cast<double *>(smem_raw);  // [(Np+1)][10] f64
    double *s_xfin = s_wtile + (size_t)(A.Np + 1) * GLG_ND;   // [28][32] f64: state in / final state out (owners <-> warp 0)
    T *s_xs = reinterpret_cast<T *>(s_xfin + GLG_NX * NL);    // [28 + 1 dummy][32] stage state in the groups' precision
    T *s_part = s_xs + (GLG_NX + 1) * NL;                     // [GLG_NSLOTS][32]: contributions, zero slot, canopy scale, lambda
    T *s_H = s_part + GLG_NSLOTS * NL;                        // [H_COUNT][32]
    T *s_C = s_H + H_COUNT * NL;                              // [C_COUNT][32] (NOISY)
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(s_xs) + GlgRoleSmem<T, NOISY>::col_bytes());
    int *s_tbl = reinterpret_cast<int *>(s_bar + 2);
    int *s_k = s_tbl + NL;
    int *s_tbl_t = s_k + NL;
    int *s_k_t = s_tbl_t + NL;
    int *s_bad = s_k_t + NL;
    int *s_misc = s_bad + NL;
    int *s_owntab = s_misc + 4;  // [NR][GLG_OWNER_TAB_WORDS]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // a CTA owns A.role_lanes (<= 32) consecutive envs; lanes beyond that are padding
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
#pragma unroll
    for (int i = 0; i < GLG_NU; ++i) u[i] = 0.0;
    double fruit_prev = 0.0;
    unsigned int ctr = 0;
    GlgColT<T, NL> Hc{s_H + lane};
    GlgColT<T, NL> Cc{s_C + lane};
    if (warp == 0) {
        double x[GLG_NX];
        if (active) {
            ctr = A.step_ctr[e];
            glg_env_prologue<NOISY>(U, A, e, wrow, ctr, Hc, Cc, x, u, d);
            fruit_prev = x[25];
        } else {
            // padding lanes integrate a copy of a benign state so every warp runs the same loop
            double d0[GLG_ND];
#pragma unroll
            for (int i = 0; i < 7; ++i) d0[i] = A.weather[i];
            glg_init_state(d0, x);
            if (NOISY) {
#pragma unroll
                for (int i = 0; i < C_COUNT; ++i) Cc[i] = (T)U.C[i];
            }
            glg_hoist(GlgConstView{U.P}, u, d0, Hc);
        }
#pragma unroll
        for (int i = 0; i < GLG_NX; ++i) {
            s_xs[i * NL + lane] = (T)x[i];
            s_xfin[i * NL + lane] = x[i];
        }
        s_bad[lane] = 0;
        s_part[GLG_SLOT_ZERO * NL + lane] = (T)0;
    }
    __syncthreads();
    if (GENERAL && warp == 1 && active) {
        // G1's general terms read the raw screen controls (its warp is 1 in both layouts); warp 0's prologue stored
        // the updated controls before the barrier above
#pragma unroll
        for (int i = 0; i < GLG_NU; ++i) u[i] = A.u[(size_t)i * A.B + e];
    }

    // ---- integration: group phase / owner phase
    T *xs_col = s_xs + lane;
    T *part_col = s_part + lane;
    const GlgXsCol<T> X{xs_col};
    double xo[NJ], acc[NJ];  // the RK4 state and stage sum stay fp64 in both precisions
    GlgOwnerRegs<NR> own;
    glg_owner_setup<NR>(U.K, warp, own);
#if GLG_NOINLINE_MASK
    if (lane == 0) glg_owner_store<NR>(own, s_owntab + warp * GLG_OWNER_TAB_WORDS);
    __syncwarp();
#endif
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int i = glg_owner_table.order[j * NR + warp];
        xo[j] = i >= 0 ? s_xfin[i * NL + lane] : 0.0;
        acc[j] = 0.0;
    }
    const double h_nom = A.dt / (double)A.n_sub;
    const double h_nom_sixth = h_nom / 6.0;
    int n_micro = 0;  // RK4 micro-steps this env executed (guarded loop only)
#ifdef GLG_PROFILE_GROUPS
    long long t_grp = 0, t_b1 = 0, t_own = 0, t_b2 = 0;
#endif
    constexpr bool GUARDED = NOISY || GRADED;
    if (!GUARDED) {
        // nominal parameters, fixed-step integrator: plain loop (the harvest guard cannot trigger: the crop approaches cLeafMax from
        // below and lambda stays ~1e-5 1/s; the guarded loop's extra state costs ~8 % at B = 4096)
        const int n_eval = 4 * A.n_sub;
#pragma unroll 1
        for (int ev = 0; ev < n_eval; ++ev) {
#ifdef GLG_PROFILE_GROUPS
            const long long c0 = clock64();
#endif
            glg_run_warp_groups<GENERAL, NOISY, false, NR, T>(warp, U, GlgKView<T>::c(U), xs_col, part_col, s_H + lane, s_C + lane, u);
#if GLG_NOINLINE_MASK
            glg_owner_fetch<NR>(own, s_owntab + warp * GLG_OWNER_TAB_WORDS);
#endif
#ifdef GLG_PROFILE_GROUPS
            const long long c1 = clock64();
#endif
            __syncthreads();
#ifdef GLG_PROFILE_GROUPS
            const long long c2 = clock64();
#endif
#if defined(GLG_PROFILE_GROUPS) || defined(GLG_PROFILE_MASK)
            if ((glg_prof_mask_dev >> 8) & 1)
#endif
            glg_owner_update<NR>(own, warp, xs_col, part_col, xo, acc, ev & 3, h_nom, h_nom_sixth);
#ifdef GLG_PROFILE_GROUPS
            const long long c3 = clock64();
#endif
            __syncthreads();
#ifdef GLG_PROFILE_GROUPS
            const long long c4 = clock64();
            t_grp += c1 - c0; t_b1 += c2 - c1; t_own += c3 - c2; t_b2 += c4 - c3;
#endif
        }
    } else {
        // Guarded loop (parametric uncertainty and / or the graded integrator): one nominal RK4 substep = m micro-steps of
        // h_nom/m, m per env = max of the harvest-stiffness guard (glg_model.h; 1 unless an organ sits inside its harvest
        // window) and, with integrator = 1, the graded start of the interval and the transient-stiffness rule.  Lanes with a
        // smaller m idle with h = 0 for the remaining micro-steps of the CTA, so an env's result never depends on its CTA mates.
        int sub = 0, q = 0, stage = 0, m_lane = 1, m_cta = 1;
        double h_lane = h_nom, h_sixth = h_nom_sixth;
#pragma unroll 1
        while (sub < A.n_sub) {
            if (NOISY) glg_run_warp_groups<GENERAL, true, true, NR, T>(warp, U, Cc, xs_col, part_col, s_H + lane, s_C + lane, u);
            else glg_run_warp_groups<GENERAL, false, true, NR, T>(warp, U, GlgKView<T>::c(U), xs_col, part_col, s_H + lane, s_C + lane, u);
#if GLG_NOINLINE_MASK
            glg_owner_fetch<NR>(own, s_owntab + warp * GLG_OWNER_TAB_WORDS);
#endif
            __syncthreads();
            if (stage == 0) {
                if (q == 0) {
                    m_lane = glg_micro_steps_from_lambda((double)part_col[GLG_SLOT_LAMBDA * NL], h_nom);
                    if (A.integrator == 1) {
                        int ms = 1 + (int)floor(h_nom * (double)part_col[GLG_SLOT_STIFF * NL] * GLG_STIFF_INV_CFL);
                        ms = ms > GLG_MAX_MICRO ? GLG_MAX_MICRO : (ms < 1 ? 1 : ms);  // ms < 1 only for a NaN estimate
                        if (sub < GLG_GRADED_SUBSTEPS && ms < GLG_GRADED_M) ms = GLG_GRADED_M;
                        m_lane = max(m_lane, ms);
                    }
                    n_micro += m_lane;
                    m_cta = __reduce_max_sync(0xffffffffu, m_lane);  // every warp sees the same 32 envs
                    // reciprocal + multiply instead of IEEE divisions: three of them (with their slow paths) pushed the loop
                    // body past the 32 KB instruction cache; the step size differs from h_nom/m by at most 1 ulp
                    h_lane = h_nom * glg_rcp((double)m_lane);
                    h_sixth = h_lane * (1.0 / 6.0);
                }
            }
            glg_owner_update<NR>(own, warp, xs_col, part_col, xo, acc, stage, q < m_lane ? h_lane : 0.0, q < m_lane ? h_sixth : 0.0);
            __syncthreads();
            if (++stage == 4) {
                stage = 0;
                if (++q >= m_cta) {
                    q = 0;
                    ++sub;
                }
            }
        }
    }
#ifdef GLG_PROFILE_GROUPS
    if (blockIdx.x == 0 && lane == 0)
        printf("warp %d: group %lld  barrier1 %lld  owner %lld  barrier2 %lld  cycles/eval\n", warp, t_grp / (4 * A.n_sub),
               t_b1 / (4 * A.n_sub), t_own / (4 * A.n_sub), t_b2 / (4 * A.n_sub));
#endif
    {
        int bad = 0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            bad |= !(fabs(xo[j]) <= 1.79769313486231570e308);
            const int i = glg_owner_table.order[j * NR + warp];
            if (i >= 0) s_xfin[i * NL + lane] = xo[j];
        }
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
            for (int i = 0; i < GLG_NX; ++i) x[i] = s_xfin[i * NL + lane];
            glg_env_epilogue(U, A, e, k, kw, tbl, wrow, x, fruit_prev, bad, ctr, o);
        }
        s_tbl[lane] = o.tbl_obs;
        s_k[lane] = o.k_obs;
        s_tbl_t[lane] = o.tbl_term;
        s_k_t[lane] = o.k_term;
        glg_stats_reduce(A, active, bad, o);
        if (GUARDED) {
            const int tot = __reduce_add_sync(0xffffffffu, active ? n_micro : 0);
            if (lane == 0 && tot > 0) atomicAdd(&A.stats[15], (double)tot);
        }
    }
    __syncthreads();
    glg_write_forecast(A, A.role_lanes, s_tbl, s_k, s_tbl_t, s_k_t, s_wtile, uniform, bk, bt);
}
