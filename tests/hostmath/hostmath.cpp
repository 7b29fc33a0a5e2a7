// Test-only host build of the device math in csrc/glg_model.h (never part of the product library).
// Lets `pytest -m "not gpu"` check the restructured RHS / RK4 against the oracle without a GPU.
#include "../../greenlight-gym2_b200/csrc/glg_model.h"
#include "../../greenlight-gym2_b200/csrc/glg_rk4.h"

extern "C" {
int hm_nominal_structure(const double *p) { return glg_params_nominal_structure(p) ? 1 : 0; }

void hm_rhs(const double *x, const double *u, const double *d, const double *p, int general, double *S) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT];
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    if (general) glg_rhs<true>(K, C, H, p, u, d, x, S);
    else glg_rhs<false>(K, C, H, p, u, d, x, S);
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
