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

// RHS assembled from the eight group functions + owner-side summation/scaling (the warp-specialised kernel's data flow)
void hm_rhs_roles(const double *x, const double *u, const double *d, const double *p, int general, double *S) {
    double K[K_COUNT], C[C_COUNT], H[H_COUNT], part[GLG_NGROUPS][GLG_NX] = {};
    glg_make_k(p, K);
    glg_make_c(p, C);
    glg_hoist(p, u, d, H);
    double *q[GLG_NGROUPS];
    for (int g = 0; g < GLG_NGROUPS; ++g) q[g] = part[g];
    double can_scale;
    if (general) {
        can_scale = glg_grp_rad<true>(K, C, H, x, q[0]);
        glg_grp_fir<true>(K, C, H, p, u, x, q[1]);
        glg_grp_conv<true>(K, C, H, p, x, q[3]);
        glg_grp_photo<true>(K, C, H, x, q[6]);
    } else {
        can_scale = glg_grp_rad<false>(K, C, H, x, q[0]);
        glg_grp_fir<false>(K, C, H, p, u, x, q[1]);
        glg_grp_conv<false>(K, C, H, p, x, q[3]);
        glg_grp_photo<false>(K, C, H, x, q[6]);
    }
    glg_grp_airflow(K, H, x, q[2]);
    glg_grp_screens(K, H, x, q[4]);
    glg_grp_cover(K, C, H, x, q[5]);
    glg_grp_flows(K, C, x, q[7]);
    for (int i = 0; i < GLG_NX; ++i) {
        double sum = 0.0;
        for (int g = 0; g < GLG_NGROUPS; ++g)
            if (glg_group_mask(i) >> g & 1u) sum += part[g][i];
        const int sk = glg_state_scale_index(i);
        S[i] = (sk >= 0 ? K[sk] : (sk == -1 ? 1.0 : can_scale)) * sum;
    }
}

// the same group functions instantiated in float (throughput mode): contributions in fp32, owner-side sum/scale in fp64
struct HmF32View { const float *b; float operator[](int i) const { return b[i]; } };
void hm_rhs_roles_f32(const double *x, const double *u, const double *d, const double *p, double *S) {
    double Kd[K_COUNT], Cd[C_COUNT], Hd[H_COUNT];
    glg_make_k(p, Kd);
    glg_make_c(p, Cd);
    glg_hoist(p, u, d, Hd);
    float Kf[K_COUNT], Cf[C_COUNT], Hf[H_COUNT], xf[GLG_NX], part[GLG_NGROUPS][GLG_NX] = {};
    for (int i = 0; i < K_COUNT; ++i) Kf[i] = (float)Kd[i];
    for (int i = 0; i < C_COUNT; ++i) Cf[i] = (float)Cd[i];
    for (int i = 0; i < H_COUNT; ++i) Hf[i] = (float)Hd[i];
    for (int i = 0; i < GLG_NX; ++i) xf[i] = (float)x[i];
    HmF32View K{Kf}, C{Cf}, H{Hf}, X{xf};
    float *q[GLG_NGROUPS];
    for (int g = 0; g < GLG_NGROUPS; ++g) q[g] = part[g];
    const float can_scale = glg_grp_rad<false>(K, C, H, X, q[0]);
    glg_grp_fir<false>(K, C, H, p, u, X, q[1]);
    glg_grp_airflow(K, H, X, q[2]);
    glg_grp_conv<false>(K, C, H, p, X, q[3]);
    glg_grp_screens(K, H, X, q[4]);
    glg_grp_cover(K, C, H, X, q[5]);
    glg_grp_photo<false>(K, C, H, X, q[6]);
    glg_grp_flows(K, C, X, q[7]);
    for (int i = 0; i < GLG_NX; ++i) {
        double sum = 0.0;
        for (int g = 0; g < GLG_NGROUPS; ++g)
            if (glg_group_mask(i) >> g & 1u) sum += (double)part[g][i];
        const int sk = glg_state_scale_index(i);
        S[i] = (sk >= 0 ? Kd[sk] : (sk == -1 ? 1.0 : (double)can_scale)) * sum;
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
