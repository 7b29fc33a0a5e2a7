// Test-only host build of the device math in csrc/glg_model.h (never part of the product library).
// Lets `pytest -m "not gpu"` check the restructured RHS / RK4 against the oracle without a GPU.
#include "../../greenlight-gym2_b200/csrc/glg_model.h"
#include "../../greenlight-gym2_b200/csrc/glg_rk4.h"

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

// RHS assembled from the four role functions + owner-side summation/scaling (the warp-specialised kernel's data flow)
void hm_rhs_roles(const double *x, const double *u, const double *d, const double *p, int general, double *S) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT], part[GLG_NROLES][GLG_NX] = {};
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    double *p0 = part[0], *p1 = part[1], *p2 = part[2], *p3 = part[3];
    if (general) glg_role_rad<true>(K, C, H, p, u, x, p0);
    else glg_role_rad<false>(K, C, H, p, u, x, p0);
    glg_role_air(K, C, H, x, p1);
    glg_role_vap(K, C, H, x, p2);
    if (general) glg_role_crop<true>(K, C, H, x, p3);
    else glg_role_crop<false>(K, C, H, x, p3);
    for (int i = 0; i < GLG_NX; ++i) {
        double sum = 0.0;
        for (int r = 0; r < GLG_NROLES; ++r)
            if (glg_role_mask(i) >> r & 1u) sum += part[r][i];
        S[i] = glg_state_scale(i, K, C, x[23]) * sum;
    }
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
