// Test-only host build of the device math in csrc/glg_model.h (never part of the product library).
// Lets `pytest -m "not gpu"` check the restructured RHS / RK4 against the oracle without a GPU.
#include "../../greenlight-gym2_b200/csrc/glg_model.h"
#include "../../greenlight-gym2_b200/csrc/glg_rk4.h"
#include "../../greenlight-gym2_b200/csrc/glg_units.h"

extern "C" {
void hm_math(int op, const double *in, double *out, int n) {
    for (int i = 0; i < n; ++i) out[i] = glg_math_eval(op, in[i]);
}
int hm_nominal_structure(const double *p) { return glg_params_nominal_structure(p) ? 1 : 0; }

void hm_rhs(const double *x, const double *u, const double *d, const double *p, int general, double *S) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT];
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    if (general) glg_rhs<true>(K, C, H, p, u, d, x, S);
    else glg_rhs<false>(K, C, H, p, u, d, x, S);
}

// RHS assembled from the flux units + owner-side summation/scaling (the warp-specialised kernel's data flow): every group warp
// of the NG-warp assignment accumulates its units into its own partial sums, the owner adds the warps' sums in warp order.
}  // extern "C"
template <class T>
struct HmView { const T *b; T operator[](int i) const { return b[i]; } };
template <class T>
struct HmX { const T *p; template <int I> T at() const { return p[I]; } };
template <int NG, int W, bool GENERAL, class T>
static void hm_warps(const HmView<T> &K, const HmView<T> &C, const HmView<T> &H, const double *p, const double *u, const HmX<T> &X,
                     double *sum, GlgSpecial<T> &sp) {
    if constexpr (W < NG) {
        T v[GLG_NX] = {};
        glg_run_warp_units<NG, W, 0, GENERAL>(K, C, H, HmView<double>{p}, u, X, v, sp);
        constexpr unsigned mask = GlgWT<NG, GENERAL>::t.states[W];
        for (int i = 0; i < GLG_NX; ++i)
            if (mask >> i & 1u) sum[i] += (double)v[i];
        hm_warps<NG, W + 1, GENERAL, T>(K, C, H, p, u, X, sum, sp);
    }
}
template <int NG, bool GENERAL, class T>
static void hm_units_rhs(const double *x, const double *u, const double *d, const double *p, double *S) {
    double Kd[K_COUNT], Cd[C_COUNT], Hd[H_COUNT];
    glg_make_k(p, Kd);
    glg_make_c(p, Cd);
    glg_hoist(p, u, d, Hd);
    T Kt[K_COUNT], Ct[C_COUNT], Ht[H_COUNT], xt[GLG_NX];
    for (int i = 0; i < K_COUNT; ++i) Kt[i] = (T)Kd[i];
    for (int i = 0; i < C_COUNT; ++i) Ct[i] = (T)Cd[i];
    for (int i = 0; i < H_COUNT; ++i) Ht[i] = (T)Hd[i];
    for (int i = 0; i < GLG_NX; ++i) xt[i] = (T)x[i];
    double sum[GLG_NX] = {};
    GlgSpecial<T> sp{};
    hm_warps<NG, 0, GENERAL, T>(HmView<T>{Kt}, HmView<T>{Ct}, HmView<T>{Ht}, p, u, HmX<T>{xt}, sum, sp);
    for (int i = 0; i < GLG_NX; ++i) {
        const int sk = glg_state_scale_index(i);
        S[i] = (sk >= 0 ? Kd[sk] : (sk == -1 ? 1.0 : (double)sp.canscale)) * sum[i];
    }
    S[28] = (double)sp.lambda;
    S[29] = glg_stiffness(HmView<double>{Kd}, (double)sp.ascr, (double)sp.avent);
}
extern "C" {
// S has 30 entries: 28 derivatives, the harvest speed and the transient-stiffness estimate
void hm_rhs_units(const double *x, const double *u, const double *d, const double *p, int general, int ng, int f32, double *S) {
    if (f32) {
        if (ng == 8) hm_units_rhs<8, false, float>(x, u, d, p, S);
        else hm_units_rhs<12, false, float>(x, u, d, p, S);
    } else if (ng == 4) {
        if (general) hm_units_rhs<4, true, double>(x, u, d, p, S);
        else hm_units_rhs<4, false, double>(x, u, d, p, S);
    } else if (ng == 8) {
        if (general) hm_units_rhs<8, true, double>(x, u, d, p, S);
        else hm_units_rhs<8, false, double>(x, u, d, p, S);
    } else if (ng == 12) {
        if (general) hm_units_rhs<12, true, double>(x, u, d, p, S);
        else hm_units_rhs<12, false, double>(x, u, d, p, S);
    } else {
        if (general) hm_units_rhs<13, true, double>(x, u, d, p, S);
        else hm_units_rhs<13, false, double>(x, u, d, p, S);
    }
}
// stiffness estimate and harvest speed as glg_rhs (kernel A) computes them, for comparison with the unit form
double hm_rhs_stiffness(const double *x, const double *u, const double *d, const double *p) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT], S[GLG_NX];
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    return glg_rhs<true>(K, C, H, p, u, d, x, S);
}

int hm_evalf(const double *x, const double *u, const double *d, const double *p, double dt, int n_sub, int general,
             double *x_next) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT], xc[GLG_NX];
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    for (int i = 0; i < GLG_NX; ++i) xc[i] = x[i];
    GlgLocalStore st;
    int bad = general ? glg_rk4_step<true>(K, C, H, p, u, d, xc, dt, n_sub, st)
                      : glg_rk4_step<false>(K, C, H, p, u, d, xc, dt, n_sub, st);
    for (int i = 0; i < GLG_NX; ++i) x_next[i] = xc[i];
    return bad;
}
}
