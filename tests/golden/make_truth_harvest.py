"""Generates tests/golden/truth_harvest_window.npz: control intervals that START ABOVE the leaf maximum -- the situation the harvest
micro-step guard exists for (a mature crop injected with set_crop_state, or parametric uncertainty redrawing cLeafMax below the
current leaf mass: BASELINE configs[2]) -- solved with scipy Radau at rtol = atol = 1e-12 on the oracle's right-hand side.
8 states spread over the rule-based season of truth_rule_based.npz, leaf mass set to 1.02 / 1.10 / 1.30 x 1.128e5 mg m-2 (the
harvest terms of aux_states.hpp:1161-1188 remove the excess at up to 5e4 mg m-2 s-1 within the first seconds of the interval).
usage: python tests/golden/make_truth_harvest.py
"""
import ctypes as C
import os
import sys

import numpy as np
from scipy.integrate import solve_ivp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import oracle_binding as ob  # noqa: E402

DP = C.POINTER(C.c_double)


def main():
    z = np.load(os.path.join(HERE, "truth_rule_based.npz"))
    lib = ob.load()
    p = np.ascontiguousarray(z["p"])
    pp = p.ctypes.data_as(DP)
    X, U, D, Y, F = [], [], [], [], []
    for i in np.linspace(0, len(z["k"]) - 1, 8).astype(int):
        for fct in (1.02, 1.10, 1.30):
            x = z["x"][i].copy()
            x[23] = fct * 1.128e5
            u, d = np.ascontiguousarray(z["u"][i]), np.ascontiguousarray(z["d"][i])
            up, dp, f = u.ctypes.data_as(DP), d.ctypes.data_as(DP), np.zeros(28)

            def rhs(t, y, up=up, dp=dp, f=f):
                y = np.ascontiguousarray(y)
                lib.glgo_rhs(y.ctypes.data_as(DP), up, dp, pp, f.ctypes.data_as(DP))
                return f.copy()
            sol = solve_ivp(rhs, (0.0, 900.0), x, method="Radau", rtol=1e-12, atol=1e-12)
            assert sol.success
            X.append(x); U.append(u); D.append(d); Y.append(sol.y[:, -1]); F.append(fct)
            print(f"point {i} leaf mass {fct:.2f} x: pruned to {sol.y[23, -1]:.1f} mg m-2", flush=True)
    np.savez_compressed(os.path.join(HERE, "truth_harvest_window.npz"), x=np.array(X), u=np.array(U), d=np.array(D), y=np.array(Y),
                        factor=np.array(F), p=p)
    print("saved", len(X), "points")


if __name__ == "__main__":
    main()
