"""Generates tests/golden/truth_rule_based.npz: tight-tolerance solutions of single control intervals taken from a full
rule-based season (BASELINE config 1: Bleiswijk GL2009 day 0, nominal parameters, RuleBasedController in the loop).

The reference integrates with CVODES at rtol = atol = 1e-6 (greenlight_model.cpp:46-63), which cannot be run here; these
vectors are what any integrator contract of this repo is measured against (tests/test_oracle_golden.py gates the default
contract at <= 1e-6 worst step).  For every selected step k the file holds the inputs (x_k, u_k, d_k) and
y = x(900 s) from scipy Radau at rtol = atol = 1e-12.

Selection (>= 200 points): every step whose controls open a screen or the vents against a colder top compartment -- the
transient-stiffness steps of SURVEY.md B.6, found as start-of-step stiffness estimate > 0.80 1/s -- every step that follows
a jump > 0.3 of a screen or the roof vents (an even sample of 80 of them), and an even sample of the quiet rest.

Right-hand side: the C oracle's glgo_rhs, which reproduces the reference's own source text bit for bit on 400 points
(rhs_golden.npz).  When /root/reference is present, a sample of the points is re-integrated with the regex translation of the
reference source (ref_translate.py) and must agree to 1e-11 (recorded in the file as `crosscheck_max`).
usage: python tests/golden/make_truth.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np
from scipy.integrate import solve_ivp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import oracle_binding as ob  # noqa: E402
from glgym.controller import RuleBasedController  # noqa: E402
from glgym.params import init_default_params  # noqa: E402
from glgym.weather import load_weather_data  # noqa: E402

DP = C.POINTER(C.c_double)


def radau(x, u, d, p, rhs):
    sol = solve_ivp(rhs, (0.0, 900.0), x, method="Radau", rtol=1e-12, atol=1e-12)
    assert sol.success
    return sol.y[:, -1]


def main():
    lib = ob.load()
    p = init_default_params().astype(np.float64)
    W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
    s29 = np.ascontiguousarray(RuleBasedController().settings_vector(), dtype=np.float64)
    # free-running season with a tight integrator (RK4 with 1800 substeps) so the visited states are the true trajectory
    env = ob.OracleEnv(W, p, ob.default_cfg(n_sub=1800))
    X, U, D, LAM = [], [], [], []
    u_prev = np.zeros(6)
    jump = []
    for k in range(5761):
        x = env.x
        u = ob.rule_control(s29, x, W[k], env.e.hour_of_day, env.e.day_of_year)
        X.append(x); U.append(u); D.append(W[k].copy()); jump.append(np.abs(u - u_prev)[[2, 3, 5]].max())
        env.step_rule(s29)
        u_prev = u
    X, U, D, jump = map(np.array, (X, U, D, jump))
    # stiffness by finite-difference Jacobian would cost 29 RHS per step: cheap enough in C
    lam = np.zeros(len(X))
    f0, f1, xp = np.zeros(28), np.zeros(28), np.zeros(28)
    for k in range(len(X)):
        J = np.zeros((28, 28))
        lib.glgo_rhs(X[k].ctypes.data_as(DP), U[k].ctypes.data_as(DP), D[k].ctypes.data_as(DP), p.ctypes.data_as(DP), f0.ctypes.data_as(DP))
        for j in range(28):
            xp[:] = X[k]
            h = 1e-6 * max(abs(xp[j]), 1e-2)
            xp[j] += h
            lib.glgo_rhs(xp.ctypes.data_as(DP), U[k].ctypes.data_as(DP), D[k].ctypes.data_as(DP), p.ctypes.data_as(DP), f1.ctypes.data_as(DP))
            J[:, j] = (f1 - f0) / h
        lam[k] = np.abs(np.linalg.eigvals(J).real).max()
    stiff = np.nonzero(lam > 0.80)[0]
    jumps = np.nonzero(jump > 0.3)[0]
    jumps = jumps[np.linspace(0, len(jumps) - 1, min(len(jumps), 80)).astype(int)]
    quiet = np.setdiff1d(np.arange(0, 5761, 96), np.concatenate([stiff, jumps]))
    sel = np.unique(np.concatenate([stiff, jumps, quiet]))
    print(f"{len(stiff)} stiff steps (lambda > 0.8, max {lam.max():.3f}), {len(jumps)} control jumps, {len(quiet)} quiet -> {len(sel)} points")
    ys = np.zeros((len(sel), 28))
    t0 = time.time()
    for i, k in enumerate(sel):
        u, d = U[k], D[k]
        up, dp, pp = u.ctypes.data_as(DP), d.ctypes.data_as(DP), p.ctypes.data_as(DP)
        f = np.zeros(28)

        def rhs(t, y, up=up, dp=dp, f=f):
            y = np.ascontiguousarray(y)
            lib.glgo_rhs(y.ctypes.data_as(DP), up, dp, pp, f.ctypes.data_as(DP))
            return f.copy()
        ys[i] = radau(X[k], u, d, p, rhs)
        if i % 40 == 0:
            print(f"  {i}/{len(sel)}  {time.time() - t0:.0f} s", flush=True)
    cross = -1.0
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, "/root/reference")
        import ref_translate as rt
        aux, ode = rt.load_reference_rhs()[:2]
        worst = 0.0
        for i in list(range(0, len(sel), max(1, len(sel) // 6)))[:6]:
            k = sel[i]
            yr = radau(X[k], U[k], D[k], p, lambda t, y, k=k: np.array(ode(list(y), list(U[k]), list(D[k]), list(p))))
            worst = max(worst, float(np.max(np.abs(yr - ys[i]) / np.maximum(np.abs(ys[i]), 1e-3))))
        cross = worst
        print(f"cross-check against the reference-translated RHS on 6 points: max rel diff {worst:.2e}")
        assert worst < 1e-11
    np.savez_compressed(os.path.join(HERE, "truth_rule_based.npz"), k=sel, x=X[sel], u=U[sel], d=D[sel], y=ys, p=p, lam=lam[sel],
                        jump=jump[sel], crosscheck_max=cross)
    print("saved", len(sel), "points")


if __name__ == "__main__":
    main()
