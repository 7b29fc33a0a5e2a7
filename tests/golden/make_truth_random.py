"""Generates tests/golden/truth_random_actions.npz: tight-tolerance solutions of single control intervals taken from seasons driven
by RANDOM actions (the distribution an exploring RL agent produces) -- the complement of truth_rule_based.npz (make_truth.py).

Two free-running seasons on the C oracle (graded RK4): one with U(-1,1) actions (what bench.py feeds), one with bang-bang actions
(+-1 per component: the largest control jumps the rate limit allows).  Every `stride`-th interval is solved with scipy Radau at
rtol = atol = 1e-12 on the oracle's right-hand side, under a per-point wall-clock budget: an interval in which the air / top-
compartment temperature difference changes sign -- the |dT|^0.66 cusp of the screen air flux, aux_states.hpp:787-814 -- makes an
error-controlled solver at 1e-12 crawl; such points are recorded as `skipped` (index and season) instead of stalling the run.
usage: python tests/golden/make_truth_random.py [stride=96] [budget_s=40]
"""
import ctypes as C
import os
import signal
import sys
import time

import numpy as np
from scipy.integrate import solve_ivp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import oracle_binding as ob  # noqa: E402
from glgym.params import init_default_params  # noqa: E402
from glgym.weather import load_weather_data  # noqa: E402

DP = C.POINTER(C.c_double)


class Budget(Exception):
    pass


def main():
    stride = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    budget = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lib = ob.load()
    p = init_default_params().astype(np.float64)
    W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
    pp = p.ctypes.data_as(DP)

    def on_alarm(signum, frame):
        raise Budget()
    signal.signal(signal.SIGALRM, on_alarm)
    X, U, D, Y, K, S, skipped = [], [], [], [], [], [], []
    t0 = time.time()
    for season, mode in enumerate(("uniform", "bang")):
        rng = np.random.default_rng(100 + season)
        env = ob.OracleEnv(W, p, ob.default_cfg(n_sub=260, stiff_guard=ob.INTEGRATOR_GRADED))
        for k in range(5760):
            a = rng.uniform(-1, 1, 6).astype(np.float32)
            if mode == "bang":
                a = np.sign(a).astype(np.float32)
            x = env.x.copy()
            env.step(action=a)
            if k % stride != stride // 2 + 7 * season:
                continue
            u, d = env.u.copy(), np.ascontiguousarray(W[k])
            up, dp = u.ctypes.data_as(DP), d.ctypes.data_as(DP)
            f = np.zeros(28)

            def rhs(t, y, up=up, dp=dp, f=f):
                y = np.ascontiguousarray(y)
                lib.glgo_rhs(y.ctypes.data_as(DP), up, dp, pp, f.ctypes.data_as(DP))
                return f.copy()
            signal.alarm(budget)
            try:
                sol = solve_ivp(rhs, (0.0, 900.0), x, method="Radau", rtol=1e-12, atol=1e-12)
                signal.alarm(0)
                assert sol.success
            except Budget:
                skipped.append((season, k))
                print(f"  season {season} step {k}: over the {budget} s budget, skipped", flush=True)
                continue
            X.append(x); U.append(u); D.append(d); Y.append(sol.y[:, -1]); K.append(k); S.append(season)
            print(f"  season {season} ({mode}) step {k}: ok, {len(X)} points, {time.time() - t0:.0f} s", flush=True)
    np.savez_compressed(os.path.join(HERE, "truth_random_actions.npz"), k=np.array(K), season=np.array(S), x=np.array(X), u=np.array(U),
                        d=np.array(D), y=np.array(Y), p=p, skipped=np.array(skipped).reshape(-1, 2))
    print("saved", len(X), "points;", len(skipped), "skipped")


if __name__ == "__main__":
    main()
