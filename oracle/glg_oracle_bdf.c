/*
 * glg_oracle_bdf.c -- CVODES-class CPU baseline for the GreenLight step.  TEST / BENCH INFRASTRUCTURE ONLY (see glg_oracle.h).
 *
 * The reference integrates one control interval with CasADi's `cvodes` plugin: SUNDIALS variable-order variable-step BDF
 * with a Newton iteration, abstol = reltol = 1e-6 (greenlight_model.cpp:46-63).  CasADi / SUNDIALS are not available in this
 * image, so the reference's solver cannot be run; this file is a plain-C adaptive implicit multistep solver of the same
 * class on the oracle's right-hand side, so that the CPU baseline bench.py reports is not handicapped by the 2400
 * right-hand-side evaluations of the fixed-step RK4(600) port:
 *   - variable order 1..5 backward differentiation formulas in backward-difference form with the NDF correction terms
 *     (Shampine & Reichelt, "The MATLAB ODE suite", SIAM J. Sci. Comput. 18, 1997 -- the published algorithm behind MATLAB's
 *     ode15s and scipy's BDF), quasi-constant step size;
 *   - simplified Newton iteration on (I - c J) with a dense LU, the LU kept while the step size is unchanged and the
 *     Jacobian kept until the iteration converges too slowly (the same reuse policy CVODES applies);
 *   - forward-difference Jacobian (28 extra right-hand sides), reused across steps and, optionally, across control intervals
 *     of the same env (a CasADi build has an exact AD Jacobian instead, so Jacobian evaluations are reported separately);
 *   - every call starts cold (order 1, fresh initial step): CasADi's integrator is re-initialised at every evalF call.
 * It is NOT used as a parity checker: the RK4 oracle is.  tests/test_oracle_golden.py checks it against the Radau(1e-12)
 * truth vectors (its error must sit inside the tolerance band it is given).
 */
#include <math.h>
#include <string.h>

#include "glg_oracle.h"

#define N GLGO_NX
#define MAX_ORDER 5
#define NEWTON_MAXITER 4
#define MIN_FACTOR 0.2
#define MAX_FACTOR 10.0

typedef struct {
    const double *u, *d, *p;
    long nfev, njev, nlu, nsteps;
} bdf_ctx;

static void fun(bdf_ctx *c, const double *y, double *f) {
    glgo_rhs(y, c->u, c->d, c->p, f);
    c->nfev++;
}

static double rms_scaled(const double *v, const double *scale) {
    double s = 0.0;
    int i;
    for (i = 0; i < N; ++i) {
        const double q = v[i] / scale[i];
        s += q * q;
    }
    return sqrt(s / N);
}

/* forward differences, column scaling like scipy's num_jac without the adaptive factor update */
static void num_jac(bdf_ctx *c, const double *y, const double *f, double *J /* [N][N] row-major */) {
    double yp[N], fp[N];
    int i, j;
    const double eps = 1.4901161193847656e-08; /* sqrt(DBL_EPSILON) */
    memcpy(yp, y, sizeof yp);
    for (j = 0; j < N; ++j) {
        double h = eps * fmax(fabs(y[j]), 1e-3);
        h = (y[j] + h) - y[j];
        yp[j] = y[j] + h;
        fun(c, yp, fp);
        for (i = 0; i < N; ++i) J[i * N + j] = (fp[i] - f[i]) / h;
        yp[j] = y[j];
    }
    c->njev++;
}

/* LU with partial pivoting of A = I - cc J ; returns 0 on success */
static int lu_factor(const double *J, double cc, double *LU, int *piv) {
    int i, j, k;
    for (i = 0; i < N; ++i)
        for (j = 0; j < N; ++j) LU[i * N + j] = (i == j ? 1.0 : 0.0) - cc * J[i * N + j];
    for (k = 0; k < N; ++k) {
        int pk = k;
        double mx = fabs(LU[k * N + k]);
        for (i = k + 1; i < N; ++i)
            if (fabs(LU[i * N + k]) > mx) {
                mx = fabs(LU[i * N + k]);
                pk = i;
            }
        if (!(mx > 0.0)) return 1;
        piv[k] = pk;
        if (pk != k)
            for (j = 0; j < N; ++j) {
                const double t = LU[k * N + j];
                LU[k * N + j] = LU[pk * N + j];
                LU[pk * N + j] = t;
            }
        for (i = k + 1; i < N; ++i) {
            const double l = LU[i * N + k] / LU[k * N + k];
            LU[i * N + k] = l;
            if (l != 0.0)
                for (j = k + 1; j < N; ++j) LU[i * N + j] -= l * LU[k * N + j];
        }
    }
    return 0;
}
static void lu_solve(const double *LU, const int *piv, double *b) {
    int i, j;
    for (i = 0; i < N; ++i) {
        const double t = b[piv[i]];
        b[piv[i]] = b[i];
        b[i] = t;
        for (j = 0; j < i; ++j) b[i] -= LU[i * N + j] * b[j];
    }
    for (i = N - 1; i >= 0; --i) {
        for (j = i + 1; j < N; ++j) b[i] -= LU[i * N + j] * b[j];
        b[i] /= LU[i * N + i];
    }
}

/* rescale the backward differences D[0..order] for a step-size change by `factor` (Shampine & Reichelt, eq. for R U) */
static void change_D(double D[MAX_ORDER + 3][N], int order, double factor) {
    double R[MAX_ORDER + 1][MAX_ORDER + 1], U[MAX_ORDER + 1][MAX_ORDER + 1], RU[MAX_ORDER + 1][MAX_ORDER + 1];
    double Dn[MAX_ORDER + 1][N];
    int i, j, k;
    for (i = 0; i <= order; ++i)
        for (j = 0; j <= order; ++j) {
            R[i][j] = 0.0;
            U[i][j] = 0.0;
        }
    /* M[i][j] = (i - 1 - factor j) / i ; R = cumprod over rows, R[0][:] = 1 */
    for (j = 0; j <= order; ++j) {
        R[0][j] = 1.0;
        U[0][j] = 1.0;
    }
    for (i = 1; i <= order; ++i)
        for (j = 1; j <= order; ++j) {
            R[i][j] = R[i - 1][j] * ((double)(i - 1) - factor * (double)j) / (double)i;
            U[i][j] = U[i - 1][j] * ((double)(i - 1) - (double)j) / (double)i;
        }
    for (i = 1; i <= order; ++i) {
        R[i][0] = 0.0;
        U[i][0] = 0.0;
    }
    for (i = 0; i <= order; ++i)
        for (j = 0; j <= order; ++j) {
            double s = 0.0;
            for (k = 0; k <= order; ++k) s += R[i][k] * U[k][j];
            RU[i][j] = s;
        }
    /* D[:order+1] = RU^T D[:order+1] */
    for (i = 0; i <= order; ++i)
        for (k = 0; k < N; ++k) {
            double s = 0.0;
            for (j = 0; j <= order; ++j) s += RU[j][i] * D[j][k];
            Dn[i][k] = s;
        }
    for (i = 0; i <= order; ++i) memcpy(D[i], Dn[i], sizeof Dn[i]);
}

/* x_next = x(dt) of x' = f(x; u, d, p), BDF/NDF orders 1..5, tolerances rtol / atol.
 * J_io: optional [28*28] Jacobian storage carried by the caller between calls (jac_valid_io says whether it holds one);
 * stats[4] (may be NULL) += {rhs evaluations, Jacobian evaluations, LU factorisations, accepted steps}.
 * returns 0, or 1 if the step failed (step size underflow / non-finite state). */
int glgo_evalf_bdf(const double *x, const double *u, const double *d, const double *p, double dt, double rtol, double atol,
                   double *x_next, double *J_io, int *jac_valid_io, long *stats) {
    static const double kappa[MAX_ORDER + 1] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
    double gamma_[MAX_ORDER + 1], alpha[MAX_ORDER + 1], error_const[MAX_ORDER + 2];
    double D[MAX_ORDER + 3][N];
    double Jloc[N * N], LU[N * N];
    double *J = J_io ? J_io : Jloc;
    int piv[N];
    bdf_ctx c;
    double t = 0.0, h, y[N], f[N], scale[N], y_predict[N], psi[N], dvec[N], dy[N], ynew[N], err[N];
    int order = 1, n_equal_steps = 0, have_lu = 0, current_jac, jac_valid, i, k, bad = 0;
    c.u = u; c.d = d; c.p = p; c.nfev = c.njev = c.nlu = c.nsteps = 0;
    gamma_[0] = 0.0;
    for (k = 1; k <= MAX_ORDER; ++k) gamma_[k] = gamma_[k - 1] + 1.0 / k;
    for (k = 0; k <= MAX_ORDER; ++k) {
        alpha[k] = (1.0 - kappa[k]) * gamma_[k];
        error_const[k] = kappa[k] * gamma_[k] + 1.0 / (k + 1);
    }
    error_const[MAX_ORDER + 1] = 1.0 / (MAX_ORDER + 2);
    memcpy(y, x, sizeof y);
    fun(&c, y, f);
    /* initial step (Hairer, Norsett & Wanner I, II.4; as in scipy's select_initial_step) */
    {
        double d0, d1, d2, h0, h1, y1[N], f1[N], tmp[N];
        for (i = 0; i < N; ++i) scale[i] = atol + rtol * fabs(y[i]);
        d0 = rms_scaled(y, scale);
        d1 = rms_scaled(f, scale);
        h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > dt) h0 = dt;
        for (i = 0; i < N; ++i) y1[i] = y[i] + h0 * f[i];
        fun(&c, y1, f1);
        for (i = 0; i < N; ++i) tmp[i] = f1[i] - f[i];
        d2 = rms_scaled(tmp, scale) / h0;
        h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 2.0);
        h = fmin(100.0 * h0, h1);
        if (h > dt) h = dt;
    }
    memset(D, 0, sizeof D);
    memcpy(D[0], y, sizeof y);
    for (i = 0; i < N; ++i) D[1][i] = f[i] * h;
    jac_valid = (J_io && jac_valid_io) ? *jac_valid_io : 0;
    if (!jac_valid) {
        num_jac(&c, y, f, J);
        jac_valid = 1;
        current_jac = 1;
    } else {
        current_jac = 0; /* carried over from the previous control interval of this env */
    }

    while (t < dt && !bad) {
        int step_accepted = 0, n_iter = 0, converged = 0;
        double safety = 1.0, error_norm = 0.0, t_new = t;
        const double min_step = 1e-12 * dt;
        while (!step_accepted) {
            double cc, dy_norm_old = -1.0, rate = -1.0;
            if (h < min_step) {
                bad = 1;
                break;
            }
            t_new = t + h;
            if (t_new > dt) {
                change_D(D, order, (dt - t) / h);
                n_equal_steps = 0;
                have_lu = 0;
                h = dt - t;
                t_new = dt;
            }
            for (i = 0; i < N; ++i) {
                double s = 0.0, ps = 0.0;
                for (k = 0; k <= order; ++k) s += D[k][i];
                for (k = 1; k <= order; ++k) ps += D[k][i] * gamma_[k];
                y_predict[i] = s;
                psi[i] = ps / alpha[order];
                scale[i] = atol + rtol * fabs(s);
            }
            cc = h / alpha[order];
            converged = 0;
            while (!converged) {
                const double tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
                if (!have_lu) {
                    if (lu_factor(J, cc, LU, piv)) {
                        bad = 1;
                        break;
                    }
                    c.nlu++;
                    have_lu = 1;
                }
                /* simplified Newton */
                memcpy(ynew, y_predict, sizeof ynew);
                memset(dvec, 0, sizeof dvec);
                dy_norm_old = -1.0;
                converged = 0;
                for (k = 0; k < NEWTON_MAXITER; ++k) {
                    double dy_norm;
                    int finite = 1;
                    fun(&c, ynew, f);
                    for (i = 0; i < N; ++i) {
                        if (!isfinite(f[i])) finite = 0;
                        dy[i] = cc * f[i] - psi[i] - dvec[i];
                    }
                    if (!finite) break;
                    lu_solve(LU, piv, dy);
                    dy_norm = rms_scaled(dy, scale);
                    rate = dy_norm_old < 0.0 ? -1.0 : dy_norm / dy_norm_old;
                    if (rate >= 0.0 && (rate >= 1.0 || pow(rate, NEWTON_MAXITER - k) / (1.0 - rate) * dy_norm > tol)) break;
                    for (i = 0; i < N; ++i) {
                        ynew[i] += dy[i];
                        dvec[i] += dy[i];
                    }
                    if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < tol)) {
                        converged = 1;
                        ++k;
                        break;
                    }
                    dy_norm_old = dy_norm;
                }
                n_iter = k;
                if (!converged) {
                    if (current_jac) break;
                    fun(&c, y_predict, f);
                    num_jac(&c, y_predict, f, J);
                    have_lu = 0;
                    current_jac = 1;
                }
            }
            if (bad) break;
            if (!converged) {
                change_D(D, order, 0.5);
                h *= 0.5;
                n_equal_steps = 0;
                have_lu = 0;
                continue;
            }
            safety = 0.9 * (2.0 * NEWTON_MAXITER + 1.0) / (2.0 * NEWTON_MAXITER + n_iter);
            for (i = 0; i < N; ++i) {
                scale[i] = atol + rtol * fabs(ynew[i]);
                err[i] = error_const[order] * dvec[i];
            }
            error_norm = rms_scaled(err, scale);
            if (error_norm > 1.0) {
                const double factor = fmax(MIN_FACTOR, safety * pow(error_norm, -1.0 / (order + 1)));
                change_D(D, order, factor);
                h *= factor;
                n_equal_steps = 0;
                have_lu = 0;
            } else {
                step_accepted = 1;
            }
        }
        if (bad) break;
        c.nsteps++;
        n_equal_steps++;
        t = t_new;
        memcpy(y, ynew, sizeof y);
        current_jac = 0; /* J was evaluated at an older point from now on */
        for (i = 0; i < N; ++i) {
            D[order + 2][i] = dvec[i] - D[order + 1][i];
            D[order + 1][i] = dvec[i];
        }
        for (k = order; k >= 0; --k)
            for (i = 0; i < N; ++i) D[k][i] += D[k + 1][i];
        if (n_equal_steps < order + 1 || t >= dt) continue;
        {
            double error_m_norm = INFINITY, error_p_norm = INFINITY, fm, f0, fp, factor;
            int delta_order = 0;
            if (order > 1) {
                for (i = 0; i < N; ++i) err[i] = error_const[order - 1] * D[order][i];
                error_m_norm = rms_scaled(err, scale);
            }
            if (order < MAX_ORDER) {
                for (i = 0; i < N; ++i) err[i] = error_const[order + 1] * D[order + 2][i];
                error_p_norm = rms_scaled(err, scale);
            }
            fm = error_m_norm > 0.0 ? pow(error_m_norm, -1.0 / order) : INFINITY;
            f0 = error_norm > 0.0 ? pow(error_norm, -1.0 / (order + 1)) : INFINITY;
            fp = error_p_norm > 0.0 ? pow(error_p_norm, -1.0 / (order + 2)) : INFINITY;
            if (!isfinite(error_m_norm)) fm = 0.0;
            if (!isfinite(error_p_norm)) fp = 0.0;
            factor = f0;
            if (fm > factor) {
                factor = fm;
                delta_order = -1;
            }
            if (fp > factor) {
                factor = fp;
                delta_order = 1;
            }
            order += delta_order;
            factor = fmin(MAX_FACTOR, safety * factor);
            change_D(D, order, factor);
            h *= factor;
            n_equal_steps = 0;
            have_lu = 0;
        }
    }
    for (i = 0; i < N; ++i) {
        x_next[i] = y[i];
        if (!isfinite(y[i])) bad = 1;
    }
    if (J_io && jac_valid_io) *jac_valid_io = bad ? 0 : 1;
    if (stats) {
        stats[0] += c.nfev;
        stats[1] += c.njev;
        stats[2] += c.nlu;
        stats[3] += c.nsteps;
    }
    return bad;
}
