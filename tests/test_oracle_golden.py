"""CPU tier: the oracle (oracle/glg_oracle.c) against the golden vectors produced from the reference
(tests/golden/make_golden.py) and against the structural expectations of the reference's own tests
(/root/reference/tests/env_test.py:17-92)."""
import os
import numpy as np
import pytest

import oracle_binding as ob
from conftest import GOLDEN, rel_err


def test_rhs_matches_reference_source(rhs_golden):
    """R1/R2: all 239 auxiliaries and 28 derivatives at 400 points, incl. perturbed crop parameters and the
    parameter sets that switch on the normally-zero terms.  Same libm, same operation order => bit-identical."""
    g = rhs_golden
    worst_a = worst_f = 0.0
    for i in range(g["x"].shape[0]):
        a, f = ob.aux_rhs(g["x"][i].copy(), g["u"][i].copy(), g["d"][i].copy(), g["p"][i].copy())
        worst_a = max(worst_a, rel_err(a, g["a"][i], 1e-300))
        worst_f = max(worst_f, rel_err(f, g["f"][i], 1e-300))
    assert worst_a <= 1e-14 and worst_f <= 1e-14, (worst_a, worst_f)


def test_rhs_survey_anchor(rhs_golden):
    """SURVEY.md Appendix E anchor: f(x0, u=0, d=W[0]) incl. the float Kelvin offset of airMv (aux_states.hpp:84)."""
    f = ob.rhs(rhs_golden["x"][0].copy(), rhs_golden["u"][0].copy(), rhs_golden["d"][0].copy(), rhs_golden["p"][0].copy())
    assert abs(f[0] - (-3.9363845622433037e-02)) < 1e-15
    assert abs(f[15] - 1.1276861623873362) < 1e-13
    assert abs(f[16] - (-1.7698364535644129e-01)) < 1e-14
    assert f[27] == 1.0 / 86400.0


def test_step_semantics_match_reference_env(shell_trace, weather0, params64):
    """S1,S3-S7: the reference's own TomatoEnv.step() trace (40 random float32 actions) reproduced by the C oracle."""
    t = shell_trace
    env = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=int(t["n_sub"])))
    assert rel_err(env.reset(), t["reset_obs"], 1e-12) <= 1e-13
    for s in range(t["step_actions"].shape[0]):
        obs, r, done, info = env.step(action=t["step_actions"][s])
        assert rel_err(obs, t["step_obs"][s], 1e-9) <= 1e-11, s
        # numpy's SIMD exp and glibc's differ in the last ulp of satVp => RH and its penalty term differ by ~1e-11
        assert abs(r - t["step_reward"][s]) <= 1e-10, s
        assert rel_err(info[[0, 1, 2, 4, 5, 6, 7, 8, 9, 10]], t["step_info"][s][[0, 1, 2, 4, 5, 6, 7, 8, 9, 10]], 1e-9) <= 1e-9, s
        assert abs(info[3] - t["step_info"][s][3]) <= 1e-15
        assert rel_err(env.x, t["step_x"][s], 1e-9) <= 1e-12, s
        assert np.array_equal(env.u, t["step_u"][s]), s
        assert done == bool(t["step_term"][s])
    assert abs(env.e.day_of_year - float(t["doy"])) < 1e-12 and abs(env.e.hour_of_day - float(t["hod"])) < 1e-12


def test_raw_control_matches_reference_env(shell_trace, weather0, params64):
    """step_raw_control (tomato_env.py:148-173) with the reference's rule-based controller outputs."""
    t = shell_trace
    env = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=int(t["n_sub"])))
    for s in range(t["rb_u"].shape[0]):
        obs, r, done, info = env.step(control=t["rb_u"][s])
        assert rel_err(obs, t["rb_obs"][s], 1e-9) <= 1e-11, s
        assert abs(r - t["rb_reward"][s]) <= 1e-10, s
    # SURVEY.md B.7 smoke anchor (first rule-based step): reward -0.2777, CO2 778.35 ppm, tAir 17.14, RH 95.6
    assert abs(t["rb_reward"][0] - (-0.27769688624)) < 1e-6 and abs(t["rb_obs"][0][0] - 778.35055) < 1e-3


def test_parametric_noise_matches_reference(shell_trace, weather0, params64):
    """S2: noise.py:3-23 applied to the float32 table; then a step() with those parameters."""
    t = shell_trace
    lib = ob.load()
    for n34, p_ref in zip(t["noise_draws"], t["noise_params"]):
        out = np.zeros(208)
        lib.glgo_param_noise(ob.P(params64), ob.P(np.ascontiguousarray(n34)), ob.P(out))
        assert np.array_equal(out, p_ref)
    env = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=int(t["n_sub"])))
    for s in range(t["noise_actions"].shape[0]):
        obs, r, done, info = env.step(action=t["noise_actions"][s], noise34=t["noise_draws"][s])
        assert rel_err(obs, t["noise_obs"][s], 1e-9) <= 1e-11, s
        assert rel_err(env.x, t["noise_x"][s], 1e-9) <= 1e-12, s


def test_reference_unit_test_expectations(weather0, params64, shell_trace):
    """What /root/reference/tests/env_test.py pins: reward normalisation constant, obs length, timestep counting,
    zero variable cost for action -1 from u=0, control bounds, episode length N+1 = 5761."""
    t = shell_trace
    assert abs(float(t["max_profit"]) - 0.328 * 900 * 1e-6 / 0.065 * 1.6) < 1e-7  # test_reward_normalisation (7 places)
    cfg = ob.default_cfg(n_sub=300)
    env = ob.OracleEnv(weather0, params64, cfg)
    obs = env.reset()
    assert obs.shape[0] == 263 and env.e.timestep == 0 and env.e.terminated == 0  # test_reset
    obs, r, done, info = env.step(action=-np.ones(6, dtype=np.float32))
    assert obs.shape[0] == 263 and env.e.timestep == 1 and isinstance(r, float)  # test_step
    assert info[2] == 0.0  # test_reward: variable_costs == 0
    assert np.all(env.u >= 0.0) and np.all(env.u <= 1.0)  # test_action_scaling
    env.step(action=np.full(6, 7.0, dtype=np.float32))
    assert np.all(env.u <= 1.0)
    # test_episode_termination: the step with timestep == N is the terminal one => 5761 steps per episode
    assert int(t["N"]) == 5760 and list(t["term_at_N"]) == [False, True]
    env.e.timestep = cfg.N - 1
    assert env.step(action=np.zeros(6, dtype=np.float32))[2] is False
    assert env.step(action=np.zeros(6, dtype=np.float32))[2] is True


def test_rk4_method_error_vs_tight_implicit_solve(params64):
    """R3: RK4(n_sub) against Radau rtol=atol=1e-12 on the reference-translated RHS (tests/golden/truth_step.npz).
    The reference's CVODES runs at 1e-6 tolerances; RK4 with the default n_sub=600 is far inside that band."""
    g = np.load(f"{GOLDEN}/truth_step.npz")
    # quiet steps sit at 1e-10..2e-9 for n_sub=600; the hardest fixture (a control jump, CO2 state) at 2.8e-7,
    # i.e. still inside the reference integrator's own 1e-6 tolerance band; the error falls with h^4.
    for n_sub, tol in ((300, 2e-6), (600, 5e-7), (900, 1e-7), (1800, 1e-8)):
        worst = 0.0
        for x, u, d, y in zip(g["x"], g["u"], g["d"], g["y"]):
            yo, bad = ob.evalf(x, u, d, g["p"], 900.0, n_sub)
            assert not bad
            worst = max(worst, rel_err(yo, y, 1e-3))
        assert worst <= tol, (n_sub, worst)


def test_rk4_unstable_step_is_flagged(weather0, params64):
    """h = 15 s is beyond the stability limit: the oracle reports a non-finite result instead of garbage."""
    from glgym.weather import init_state
    y, bad = ob.evalf(init_state(weather0[0]), np.zeros(6), weather0[0], params64, 900.0, 60)
    assert bad == 1


def test_batch_evalf_threads_agree(rhs_golden, params64):
    g = rhs_golden
    sel = [i for i in range(0, 64) if i % 3 != 2 and i % 4 != 3][:16]
    x, u, d = g["x"][sel], g["u"][sel], g["d"][sel]
    y1 = ob.evalf_batch(x, u, d, params64, n_sub=300, n_threads=1)
    y4 = ob.evalf_batch(x, u, d, params64, n_sub=300, n_threads=4)
    assert np.array_equal(y1, y4)
    assert np.array_equal(y1[3], ob.evalf(x[3], u[3], d[3], params64, 900.0, 300)[0])


def test_harvest_stiffness_guard(weather0, params64):
    """Perturbed cLeafMax below the current leaf mass (noise.py:16-22 can do that at uncertainty 0.3): the harvest sigmoid's
    rate constant reaches ~10 1/s.  With the micro-step guard RK4(600) stays close to a 40x finer integration; the guard
    is inert (m = 1) for the nominal table, where results equal the unguarded scheme bit for bit."""
    from glgym.weather import init_state
    lib = ob.load()
    x0, d = init_state(weather0[0]), weather0[0].copy()
    n34 = np.zeros(34)
    n34[141 - 128], n34[142 - 128] = -0.15, 0.15  # laiMax down, sla up => cLeafMax' = 83.4e3 < cLeaf = 95.3e3
    p = np.zeros(208)
    lib.glgo_param_noise(ob.P(params64), ob.P(n34), ob.P(p))
    assert p[144] < x0[23] - 5e3
    y600, bad = ob.evalf(x0, np.zeros(6), d, p, 900.0, 600)
    yfine, _ = ob.evalf(x0, np.zeros(6), d, p, 900.0, 24000)
    assert not bad and rel_err(y600, yfine) <= 2e-4, rel_err(y600, yfine)
    assert abs(y600[23] - p[144]) < 2.0e4 and y600[23] < x0[23]  # pruned towards the new maximum, no overshoot to ~2e4
    # nominal table: the guard never splits a substep => identical to plain RK4 written out here
    x, h = x0.copy(), 900.0 / 300
    for _ in range(300):
        k1 = ob.rhs(x, np.zeros(6), d, params64); k2 = ob.rhs(x + 0.5 * h * k1, np.zeros(6), d, params64)
        k3 = ob.rhs(x + 0.5 * h * k2, np.zeros(6), d, params64); k4 = ob.rhs(x + h * k3, np.zeros(6), d, params64)
        acc = k1.copy(); acc += 2.0 * k2; acc += 2.0 * k3
        x = x + (h / 6.0) * (acc + k4)
    assert np.array_equal(x, ob.evalf(x0, np.zeros(6), d, params64, 900.0, 300)[0])


def test_graded_integrator_is_cheaper_and_more_accurate(weather0, params64):
    """Graded RK4 (glgo_evalf_ex, stiff_guard = 3, n_sub = 260): the first 12 nominal substeps of a control interval are split
    in 16, 8, 4 x4, 2 x6 and any substep is split further while the transient-stiffness estimate asks for it.  Along a
    rule-based episode prefix that contains the screen-opening transients of SURVEY B.6 (where fixed-step RK4(300) diverges),
    measured against RK4(2400): never worse than 1e-7, at least 100x more accurate than the fixed 600-substep contract in the
    worst step, with ~300 instead of 600 micro-steps."""
    import oracle_binding as ob
    from glgym.controller import RuleBasedController
    s29 = RuleBasedController().settings_vector()
    env = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=600))
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3))
    e_fixed, e_graded, micro, diverged300 = [], [], [], 0
    for k in range(360):
        x0, d = env.x.copy(), weather0[k].copy()
        u = ob.rule_control(s29, x0, d, env.e.hour_of_day, env.e.day_of_year)
        if k >= 300 or k % 15 == 0:
            ref, _, _ = ob.evalf_ex(x0, u, d, params64, 900.0, 2400, 0)
            y6, _, _ = ob.evalf_ex(x0, u, d, params64, 900.0, 600, 0)
            yg, bad, n = ob.evalf_ex(x0, u, d, params64, 900.0, 260, 3)
            y3, bad3, _ = ob.evalf_ex(x0, u, d, params64, 900.0, 300, 0)
            assert not bad
            diverged300 += int(bad3 or not np.all(np.isfinite(y3)))
            e_fixed.append(rel(y6, ref)); e_graded.append(rel(yg, ref)); micro.append(n)
        env.step(control=u)
    e_fixed, e_graded, micro = np.array(e_fixed), np.array(e_graded), np.array(micro)
    assert diverged300 >= 1                      # the prefix really contains a step fixed RK4(300) cannot do
    assert e_graded.max() <= 1e-7 and e_graded.max() <= e_fixed.max() / 100
    assert np.median(e_graded) <= 2 * np.median(e_fixed) + 1e-10
    assert 300 <= micro.mean() <= 315 and micro.max() <= 400
    # flag 0 is the fixed-step contract, bit for bit
    y_a, _, n_a = ob.evalf_ex(x0, u, d, params64, 900.0, 600, 0)
    y_b, _ = ob.evalf(x0, u, d, params64, 900.0, 600)
    assert np.array_equal(y_a, y_b) and n_a == 600


# ------------------------------------------------------------------------------------------------ integrator contract
@pytest.fixture(scope="module")
def truth_rb():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "truth_rule_based.npz"))


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))


def test_default_integrator_contract_against_tight_truth(truth_rb):
    """The integrator contract is decided on accuracy: 249 control intervals of a rule-based season --
    every transient-stiffness step (lambda_max up to 1.497 1/s), 80 screen / vent jumps, 60 quiet steps -- each solved
    with Radau at rtol = atol = 1e-12 (tests/golden/make_truth.py; the reference's CVODES runs at 1e-6 and is not available).
    Gate: the DEFAULT contract (graded RK4, n_sub = 260, 300 RK4 steps) is within 1e-6 of truth in EVERY step (measured 5.2e-8;
    the 349-step grid of n_sub = 300 with a 15-substep graded start measured 2.7e-8); the equal-substep RK4(600) grid is not
    (8.9e-5), which is why it is no longer the default."""
    import oracle_binding as ob
    z = truth_rb
    n = len(z["k"])
    assert n >= 200 and float(z["lam"].max()) > 1.4 and float(z["crosscheck_max"]) <= 1e-11
    eg, ef, micro = np.zeros(n), np.zeros(n), np.zeros(n)
    for i in range(n):
        yg, bad, m = ob.evalf_ex(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, 260, 3)
        assert not bad
        eg[i], micro[i] = _rel(yg, z["y"][i]), m
        ef[i] = _rel(ob.evalf(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, 600)[0], z["y"][i])
    assert eg.max() <= 1e-6                      # the gate
    assert eg.max() <= 1e-7 and np.percentile(eg, 99) <= 5e-8 and np.median(eg) <= 2e-9   # what is measured (regression guard)
    assert ef.max() > 1e-6 and eg.max() <= ef.max() / 1000
    assert micro.max() <= 1.2 * 300 and micro.min() >= 300 and micro.mean() <= 301
    from glgym.vec_env import DEFAULT_INTEGRATOR
    assert DEFAULT_INTEGRATOR == "graded"


def test_default_integrator_contract_on_random_action_intervals():
    """The same gate on the action distribution an exploring agent produces: 120 control intervals of two free-running seasons --
    U(-1,1) actions (what bench.py feeds) and bang-bang actions (the largest control jumps the rate limit allows) -- each solved with
    Radau at rtol = atol = 1e-12 (tests/golden/make_truth_random.py).  Default contract: <= 1e-6 in every interval (measured 4.4e-10
    uniform, 1.3e-9 bang-bang), exactly 300 RK4 steps each."""
    import oracle_binding as ob
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "truth_random_actions.npz"))
    n = len(z["k"])
    assert n >= 100 and set(z["season"].tolist()) == {0, 1} and len(z["skipped"]) == 0
    eg, micro = np.zeros(n), np.zeros(n)
    for i in range(n):
        yg, bad, m = ob.evalf_ex(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, 260, 3)
        assert not bad
        eg[i], micro[i] = _rel(yg, z["y"][i]), m
    assert eg.max() <= 1e-6                       # the gate
    assert eg.max() <= 1e-8 and np.median(eg) <= 1e-10   # what is measured (regression guard)
    assert micro.min() == 300 and micro.max() <= 320


def test_harvest_guard_against_tight_truth():
    """The harvest micro-step guard is accurate, not only self-consistent: 24 control intervals that start 2 - 30 % above the leaf
    maximum (mature crop injected / cLeafMax redrawn by parametric uncertainty), each solved with Radau at rtol = atol = 1e-12
    (tests/golden/make_truth_harvest.py).  Default contract: <= 1e-6 (measured 4.8e-9) with 568 - 603 RK4 steps; without the guard a
    plain substep overshoots the maximum by 3e4 mg m-2."""
    import oracle_binding as ob
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "truth_harvest_window.npz"))
    n = len(z["factor"])
    assert n >= 24
    for i in range(n):
        yg, bad, m = ob.evalf_ex(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, 260, 3)
        assert not bad and 500 <= m <= 700, m
        assert z["x"][i][23] - yg[23] > 1e4          # the excess leaf mass really was pruned inside the interval
        assert _rel(yg, z["y"][i]) <= 2e-8, (i, _rel(yg, z["y"][i]))   # gate 1e-6; regression guard at what is measured


def test_implicit_cpu_baseline_solver(truth_rb):
    """oracle/glg_oracle_bdf.c -- the CVODES-class CPU baseline (variable-order BDF/NDF, rtol = atol = 1e-6): its error against
    truth sits in the band an implicit multistep solver at that tolerance delivers (2e-7 ... 2e-6 per step in quiet
    steps, more in transients), with a few hundred right-hand sides per interval instead of RK4(600)'s 2400; tightening the
    tolerance tightens the result (it is a convergent solver, not a tuned one)."""
    import oracle_binding as ob
    z = truth_rb
    idx = range(0, len(z["k"]), 3)
    e6, rhs = [], []
    for i in idx:
        y, bad, st = ob.evalf_bdf(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, 1e-6, 1e-6)
        assert not bad
        e6.append(_rel(y, z["y"][i])); rhs.append(st["rhs"])
    e6, rhs = np.array(e6), np.array(rhs)
    assert e6.max() <= 1e-4 and np.median(e6) <= 5e-6
    assert rhs.mean() <= 600 and rhs.max() <= 1500
    i = int(np.argmax(z["lam"]))
    errs = [_rel(ob.evalf_bdf(z["x"][i], z["u"][i], z["d"][i], z["p"], 900.0, tol, tol)[0], z["y"][i]) for tol in (1e-5, 1e-7, 1e-9)]
    assert errs[2] < errs[1] < errs[0] and errs[2] <= 1e-7
    # through the env: same step semantics, integrator swapped (glgo_env_cfg.stiff_guard = 16)
    from glgym.weather import load_weather_data
    W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
    ea = ob.OracleEnv(W, z["p"], ob.default_cfg(n_sub=600))
    eb = ob.OracleEnv(W, z["p"], ob.default_cfg(stiff_guard=ob.INTEGRATOR_BDF))
    rng = np.random.default_rng(3)
    for s in range(5):
        a = rng.uniform(-1, 1, 6).astype(np.float32)
        oa, ra, _, _ = ea.step(action=a)
        ob_, rb, _, _ = eb.step(action=a)
        assert abs(ra - rb) <= 1e-5 and _rel(eb.x, ea.x) <= 1e-4
        assert 50 <= eb.e.n_micro <= 1500


def test_observation_module_stacks_match_reference_env():
    """Non-default observation stacks (tomato_env.py:77-96,193-198): the oracle's rows and rewards against the reference's own
    env run with those stacks (tests/golden/make_golden.py obs -> shell_trace_obs.npz); names and Box bounds of the package
    against the reference's."""
    import oracle_binding as ob
    from glgym.params import init_default_params
    from glgym.vec_env import OBSERVATION_MODULES, obs_names
    from glgym.weather import load_weather_data
    t = np.load(os.path.join(os.path.dirname(__file__), "golden", "shell_trace_obs.npz"))
    W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
    p = init_default_params().astype(np.float64)
    for si in range(int(t["n_stacks"])):
        mods = [str(m) for m in t[f"mods{si}"]]
        env = ob.OracleEnv(W, p, ob.default_cfg(n_sub=int(t["n_sub"]), obs_modules=mods))
        o0 = env.reset()
        assert o0.shape == t[f"reset_obs{si}"].shape and np.array_equal(o0, t[f"reset_obs{si}"])
        for s in range(t[f"actions{si}"].shape[0]):
            o, r, dn, _ = env.step(action=t[f"actions{si}"][s])
            assert np.allclose(o, t[f"obs{si}"][s], rtol=1e-13, atol=1e-13) and abs(r - t[f"reward{si}"][s]) <= 1e-12, (si, s)
        assert obs_names(48, mods) == [str(n) for n in t[f"names{si}"]]
        sizes = [len(OBSERVATION_MODULES[m][1]) if OBSERVATION_MODULES[m][1] is not None else 240 for m in mods]
        low = np.concatenate([np.full(n, OBSERVATION_MODULES[m][2][0]) for m, n in zip(mods, sizes)]).astype(np.float32)
        high = np.concatenate([np.full(n, OBSERVATION_MODULES[m][2][1]) for m, n in zip(mods, sizes)]).astype(np.float32)
        assert np.array_equal(low, t[f"low{si}"]) and np.array_equal(high, t[f"high{si}"])
