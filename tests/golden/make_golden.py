"""Generates the golden fixtures under tests/golden/ FROM THE REFERENCE (runs only in the build container where
/root/reference exists; the GPU box and the test-suite only read the committed .npz/.npy outputs).

  rhs_golden.npz    : 400 (x,u,d,p) points with the 239 auxiliaries and 28 derivatives obtained by executing the
                      reference's own model source (aux_states.hpp update(), ode.hpp ODE()) through ref_translate.py.
                      Pins the oracle's RHS (R1, R2).
  params_numpy2.npy : the reference's init_default_params(208) as produced under this container's numpy 2.x.
  weather_golden.npz: reference load_weather_data(Bleiswijk, GL, 2009, start_day) checksums for start days 0..18 and the
                      first/last rows + a strided sample of the start-day-0 table; init_state(W[0]).
  shell_trace.npz   : the reference's own TomatoEnv / GreenhouseReward / observation modules / noise / rule-based
                      controller executed unchanged (stub `gymnasium`, stub native module backed by the oracle's evalF):
                      (a) 40 steps of step() with seeded random actions, (b) 40 steps of the rule-based controller via
                      step_raw_control(), (c) 12 steps with uncertainty_scale=0.3 recording the perturbed parameters.
                      Pins the step semantics S1..S8 (obs, reward, info, termination, time, noise application).
                      numpy-1.26 promotion (the reference's pinned version) is emulated: env.p stays float32 (legacy
                      table), the reward reads it widened to float64 (see make_env).
  shell_trace_obs.npz: the same reference env with four non-default observation stacks (module subsets / orders), 12 steps each.
  truth_step.npz    : x(900 s) for 6 (x,u,d) points from scipy Radau rtol=atol=1e-12 on the reference-translated RHS:
                      bounds the method error of RK4(n_sub) (the reference's CVODES itself is not available).
usage: python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")
REF_WEATHER = "/root/reference/gl_gym/environments/weather"

import ref_translate as rt  # noqa: E402
import oracle_binding as ob  # noqa: E402


def install_stubs(n_sub):
    """gymnasium + native-module stubs so the reference's env code imports unchanged."""
    gym = types.ModuleType("gymnasium")

    class Env:
        def reset(self, seed=None, options=None):
            if seed is not None or not hasattr(self, "_np_random"):
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = shape if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
            self.shape, self.dtype = tuple(shape), dtype

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = Box
    spaces.Dict = dict
    utils = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))), seed)
    utils.seeding = seeding
    gym.Env, gym.spaces, gym.utils = Env, spaces, utils
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.utils": utils,
                        "gymnasium.utils.seeding": seeding})
    native = types.ModuleType("gl_gym.environments.models.greenlight_model")

    class GreenLight:
        def __init__(self, nx, nu, nd, np_, dt):
            self.dt = float(dt)

        def evalF(self, x, u, d, p):
            y, bad = ob.evalf(np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64),
                              np.asarray(d, dtype=np.float64), np.asarray(p, dtype=np.float64), self.dt, n_sub)
            if bad:
                raise RuntimeError("non-finite")
            return list(y)

    native.GreenLight = GreenLight
    sys.modules["gl_gym.environments.models.greenlight_model"] = native


def make_env(uncertainty_scale=0.0, mods=None):
    from gl_gym.environments.tomato_env import TomatoEnv
    base = dict(weather_data_dir=REF_WEATHER, location="Bleiswijk", data_source="GL", num_params=208, nx=28, nu=6, nd=10,
                dt=900, u_min=[0] * 6, u_max=[1] * 6, delta_u_max=0.1, pred_horizon=0.5, season_length=60,
                start_train_year=2009, end_train_year=2009, start_train_day=0, end_train_day=0, training=True)
    con = dict(co2_min=300., co2_max=1600., temp_min=15., temp_max=34., rh_min=50., rh_max=85.)
    rp = dict(fixed_greenhouse_cost=15., fixed_co2_cost=0.015, fixed_lamp_cost=0.07, fixed_screen_cost=2., elec_price=0.3,
              heating_price=0.09, co2_price=0.3, fruit_price=1.6, dmfm=0.065, pen_weights=[4.e-4, 5.e-3, 7.e-4], pen_lamp=0.1)
    mods = mods or ["IndoorClimateObservations", "BasicCropObservations", "ControlObservations", "WeatherObservations",
                    "TimeObservations", "WeatherForecastObservations"]
    env = TomatoEnv("GreenhouseReward", mods, con, dict(eval_days=[0], eval_years=[2009], location="Bleiswijk", data_source="GL"),
                    rp, base, uncertainty_scale)
    # numpy-1.26 emulation (the reference's pinned numpy): env.p stays a float32 array like in the reference, holding
    # the legacy-promotion table (p[169], p[171]); the reward, which under numpy 1.26 evaluates
    # `np.float32 * python_float` in float64, sees the same values widened to float64 through a proxy.
    sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
    from glgym.params import init_default_params
    env.p = init_default_params(208)

    class RewardEnvProxy:
        def __init__(self, e):
            object.__setattr__(self, "_e", e)

        def __getattr__(self, name):
            v = getattr(self._e, name)
            return v.astype(np.float64) if name == "p" else v

    env.reward.env = RewardEnvProxy(env)
    env.reward.max_profit = env.reward.max_profit_reward()
    env.reward.min_profit = env.reward.min_profit_reward()
    return env


OBS_STACKS = [  # non-default observation stacks (tomato_env.py:77-96): ablations an RL user would try
    ["IndoorClimateObservations", "BasicCropObservations", "ControlObservations", "WeatherObservations", "TimeObservations"],
    ["TimeObservations", "ControlObservations", "IndoorClimateObservations"],
    ["WeatherForecastObservations", "BasicCropObservations", "IndoorClimateObservations"],
    ["WeatherObservations", "IndoorClimateObservations", "WeatherForecastObservations", "TimeObservations"],
]


def obs_stack_traces():
    """shell_trace_obs.npz: the reference's own env with non-default observation stacks, 12 steps of seeded random actions
    each (reset observation, observations, rewards -- the reward reads obs[0:3] whatever the stack puts there)."""
    NSUB = 300
    install_stubs(NSUB)
    out = dict(n_sub=NSUB, n_stacks=len(OBS_STACKS))
    arng = np.random.default_rng(11)
    for si, mods in enumerate(OBS_STACKS):
        env = make_env(0.0, mods)
        o0, _ = env.reset(seed=666)
        acts, obs_l, rew_l = [], [], []
        for s in range(12):
            a = arng.uniform(-1, 1, 6).astype(np.float32)
            o, r, term, trunc, info = env.step(a)
            acts.append(a); obs_l.append(np.asarray(o, dtype=np.float64)); rew_l.append(float(r))
        assert env.get_obs_names() and len(env.get_obs_names()) == len(o0)
        out.update({f"mods{si}": np.array(mods), f"reset_obs{si}": np.asarray(o0, dtype=np.float64), f"actions{si}": np.array(acts),
                    f"obs{si}": np.array(obs_l), f"reward{si}": np.array(rew_l), f"names{si}": np.array(env.get_obs_names()),
                    f"low{si}": env.observation_space.low, f"high{si}": env.observation_space.high})
    np.savez_compressed(os.path.join(HERE, "shell_trace_obs.npz"), **out)
    print("shell_trace_obs.npz written:", [len(out[f"reset_obs{i}"]) for i in range(len(OBS_STACKS))])


def main():
    rng = np.random.default_rng(20261017)
    from gl_gym.environments.parameters import init_default_params
    from gl_gym.environments.utils import load_weather_data, init_state
    p_ref = init_default_params(208)
    np.save(os.path.join(HERE, "params_numpy2.npy"), p_ref)
    p = p_ref.astype(np.float64)
    W = load_weather_data(REF_WEATHER, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
    x0 = init_state(W[0])

    # ---- RHS goldens from the reference source text
    # the 12 helper functions are translated from the header text like the model body (ref_translate.translate_helpers)
    aux, ode, na, no = rt.load_reference_rhs()
    assert (na, no) == (239, 28)
    N = 400
    X, U, D, P, A, F = (np.zeros((N, n)) for n in (28, 6, 10, 208, 239, 28))
    for t in range(N):
        d = W[rng.integers(0, len(W))].copy()
        x = x0.copy()
        x[0:2] *= rng.uniform(0.5, 2.5, 2); x[2:15] += rng.uniform(-12, 14, 13); x[15:17] *= rng.uniform(0.4, 1.3, 2)
        x[17:22] += rng.uniform(-10, 30, 5); x[22] = rng.uniform(0, 25000); x[23] *= rng.uniform(0.3, 1.3)
        x[24] *= rng.uniform(0.5, 2); x[25] *= rng.uniform(0.5, 60); x[26] = rng.uniform(-50, 4000); x[27] = rng.uniform(0, 60)
        u = rng.uniform(0, 1, 6)
        if t % 5 == 0: u[rng.integers(0, 6)] = 0.0
        if t % 7 == 0: u[rng.integers(0, 6)] = 1.0
        if t == 0: x, u, d = x0.copy(), np.zeros(6), W[0].copy()      # SURVEY Appendix E anchor point
        if t == 1: x, u, d = x0.copy(), np.full(6, 0.1), W[0].copy()
        pp = p.copy()
        if t % 3 == 2:
            pp[128:162] = (pp[128:162] * (1 + rng.uniform(-0.15, 0.15, 34))).astype(np.float32)
            pp[144] = np.float32(pp[141]) / np.float32(pp[142])
        if t % 4 == 3:  # exercise the terms that vanish for the default table (GENERAL kernel variant)
            pp[70], pp[194], pp[195], pp[165], pp[198], pp[33] = 0.05, 0.03, 0.9, 0.5, 2.0, 0.65
        X[t], U[t], D[t], P[t] = x, u, d, pp
        A[t] = aux(list(x), list(u), list(d), list(pp))
        F[t] = ode(list(x), list(u), list(d), list(pp))
    np.savez_compressed(os.path.join(HERE, "rhs_golden.npz"), x=X, u=U, d=D, p=P, a=A, f=F)

    # ---- weather goldens
    sums, firsts, lasts = [], [], []
    for sd in range(19):
        Ws = load_weather_data(REF_WEATHER, "Bleiswijk", "GL", 2009, sd, 60, 49, 900, 10)
        sums.append([Ws.sum(), (Ws * np.arange(1, 11)).sum(), np.abs(Ws).max()])
        firsts.append(Ws[0]); lasts.append(Ws[-1])
    np.savez_compressed(os.path.join(HERE, "weather_golden.npz"), sums=np.array(sums), first=np.array(firsts),
                        last=np.array(lasts), sample_sd0=W[::97], shape=np.array(W.shape), x0=x0)

    # ---- truth steps (method error of RK4)
    from scipy.integrate import solve_ivp
    xs, us, ds, ys = [], [], [], []
    pts = [(x0, np.zeros(6), W[0])]
    xx = x0.copy()
    for k in range(1, 6):
        u = rng.uniform(0, 1, 6)
        xx, _ = ob.evalf(xx, u, W[40 * k], p, 900.0, 900)
        pts.append((xx.copy(), rng.uniform(0, 1, 6), W[40 * k + 1]))
    for (x, u, d) in pts:
        sol = solve_ivp(lambda t, y: np.array(ode(list(y), list(u), list(d), list(p))), (0, 900.0), x, method="Radau",
                        rtol=1e-12, atol=1e-12)
        xs.append(x); us.append(u); ds.append(d); ys.append(sol.y[:, -1])
    np.savez_compressed(os.path.join(HERE, "truth_step.npz"), x=np.array(xs), u=np.array(us), d=np.array(ds), y=np.array(ys), p=p)

    # ---- reference-shell traces (n_sub = 300, h = 3 s: stable for these steps and fast to generate; the semantics do not depend on n_sub)
    NSUB = 300
    install_stubs(NSUB)
    env = make_env(0.0)
    obs0, _ = env.reset(seed=666)
    tr = dict(n_sub=NSUB, reset_obs=np.asarray(obs0, dtype=np.float64), N=env.N, Np=env.Np,
              max_profit=float(env.reward.max_profit), min_profit=float(env.reward.min_profit),
              fixed_costs=float(env.reward.fixed_costs))
    acts, obs_l, rew_l, info_l, x_l, u_l, term_l = [], [], [], [], [], [], []
    arng = np.random.default_rng(7)
    for s in range(40):
        a = arng.uniform(-1, 1, 6).astype(np.float32)
        o, r, term, trunc, info = env.step(a)
        acts.append(a); obs_l.append(np.asarray(o, dtype=np.float64)); rew_l.append(float(r)); term_l.append(bool(term))
        info_l.append([float(info[k]) for k in ("EPI", "revenue", "variable_costs", "fixed_costs", "co2_cost", "heat_cost",
                                                "elec_cost", "temp_violation", "co2_violation", "rh_violation", "lamp_violation")])
        x_l.append(np.asarray(env.x, dtype=np.float64)); u_l.append(np.asarray(env.u, dtype=np.float64))
    tr.update(step_actions=np.array(acts), step_obs=np.array(obs_l), step_reward=np.array(rew_l), step_info=np.array(info_l),
              step_x=np.array(x_l), step_u=np.array(u_l), step_term=np.array(term_l),
              doy=env.day_of_year, hod=env.hour_of_day)
    # termination: jump to the end of the season
    env.timestep = env.N - 1
    t_l = []
    for s in range(2):
        o, r, term, trunc, info = env.step(np.zeros(6, dtype=np.float32))
        t_l.append(bool(term))
    tr["term_at_N"] = np.array(t_l)

    # rule-based controller trace
    import yaml
    from gl_gym.environments.baseline import RuleBasedController
    ctrl = RuleBasedController(**yaml.safe_load(open("/root/reference/gl_gym/configs/agents/rule_based.yml"))["TomatoEnv"])
    env = make_env(0.0)
    env.reset(seed=666)
    cu, co, cr, cx = [], [], [], []
    for s in range(40):
        u = ctrl.predict(env.x, env.weather_data[env.timestep], env)
        o, r, term, trunc, info = env.step_raw_control(u)
        cu.append(np.asarray(u, dtype=np.float64)); co.append(np.asarray(o, dtype=np.float64)); cr.append(float(r))
        cx.append(np.asarray(env.x, dtype=np.float64))
    tr.update(rb_u=np.array(cu), rb_obs=np.array(co), rb_reward=np.array(cr), rb_x=np.array(cx))

    # parametric uncertainty: record the perturbed parameter vector the reference hands to evalF
    import gl_gym.environments.tomato_env as te
    recorded = []
    orig_noise = te.parametric_crop_uncertainty

    def wrapped(parameters, uncertainty, RNG):
        state = RNG.bit_generator.state
        draws = np.random.Generator(np.random.PCG64()).uniform  # placeholder to keep flake quiet
        del draws
        RNG.bit_generator.state = state
        noise = RNG.uniform(-uncertainty / 2, uncertainty / 2, size=34)
        RNG.bit_generator.state = state
        out = orig_noise(parameters, uncertainty, RNG)  # parameters is the float32 table, as in the reference
        recorded.append((noise, np.asarray(out, dtype=np.float64)))
        return out

    te.parametric_crop_uncertainty = wrapped
    env = make_env(0.3)
    env.reset(seed=5)
    na, no, nr, nx = [], [], [], []
    for s in range(12):
        a = arng.uniform(-1, 1, 6).astype(np.float32)
        o, r, term, trunc, info = env.step(a)
        na.append(a); no.append(np.asarray(o, dtype=np.float64)); nr.append(float(r)); nx.append(np.asarray(env.x, dtype=np.float64))
    te.parametric_crop_uncertainty = orig_noise
    tr.update(noise_actions=np.array(na), noise_obs=np.array(no), noise_reward=np.array(nr), noise_x=np.array(nx),
              noise_draws=np.array([n for n, _ in recorded]), noise_params=np.array([q for _, q in recorded]))
    np.savez_compressed(os.path.join(HERE, "shell_trace.npz"), **tr)

    # rule-based controller known-answer vectors: the reference's RuleBasedController.predict (baseline.py:68-227) on
    # random (x, d, clock) points, for the shipped settings and for variants that exercise the wrapping lamp window,
    # the day-of-year window and the lamps_on == lamps_off case
    import types
    base = yaml.safe_load(open("/root/reference/gl_gym/configs/agents/rule_based.yml"))["TomatoEnv"]
    names = list(base.keys())
    variants = [dict(base), dict(base, lamps_on=20, lamps_off=6), dict(base, lamps_day_start=300, lamps_day_stop=60),
                dict(base, lamps_on=7, lamps_off=7), dict(base, lamps_on=2, lamps_off=20, heat_correction=1.5, useBlScr=0,
                                                          lamp_rad_sum_limit=4, lamps_off_sun=150)]
    krng = np.random.default_rng(2024)
    K = 160
    ks, kx, kd, kh, kdoy, ku = [], [], [], [], [], []
    for vi, var in enumerate(variants):
        ctrl_v = RuleBasedController(**var)
        for _ in range(K):
            x = np.array(tr["rb_x"][krng.integers(0, 40)])
            x[0] = krng.uniform(400, 2500); x[2] = krng.uniform(5, 35); x[15] = krng.uniform(300, 3500)
            d = np.array([krng.choice([0.0, krng.uniform(0, 800)]), krng.uniform(-10, 30), krng.uniform(200, 2000),
                          krng.uniform(700, 800), krng.uniform(0, 15), krng.uniform(-30, 15), krng.uniform(5, 15),
                          krng.uniform(0, 20), krng.choice([0.0, 1.0, krng.uniform(0, 1)]), krng.choice([0.0, 1.0, krng.uniform(0, 1)])])
            fake = types.SimpleNamespace(nu=6, hour_of_day=float(krng.choice([krng.uniform(0, 24), float(krng.integers(0, 24))])),
                                         day_of_year=float(krng.uniform(0, 365)))
            with np.errstate(all="ignore"):
                u = ctrl_v.predict(x, d, fake)
            ks.append([float(var[n]) for n in names]); kx.append(x); kd.append(d); kh.append(fake.hour_of_day)
            kdoy.append(fake.day_of_year); ku.append(np.asarray(u, dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, "ctrl_golden.npz"), names=np.array(names), settings=np.array(ks), x=np.array(kx),
                        d=np.array(kd), hod=np.array(kh), doy=np.array(kdoy), u=np.array(ku))
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith((".npz", ".npy"))))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "obs":
        obs_stack_traces()  # only the observation-stack traces (the other fixtures stay as committed)
    else:
        main()
        obs_stack_traces()
