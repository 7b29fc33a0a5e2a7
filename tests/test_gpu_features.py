"""GPU tier, round 2: observation-module stacks, the single-env TomatoEnv facade (the reference's own unit tests restated),
checkpoint / restore of the full env state, the harvest-window guard in every kernel, the headline kernel against the oracle at
the full BASELINE batch, the NCCL statistics all-reduce entry points.  All through the C-ABI (ctypes)."""
import os

import numpy as np
import pytest
import torch

import oracle_binding as ob
import philox_ref
from conftest import rel_err

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_env(B, **kw):
    from glgym.vec_env import GreenLightVecEnv
    kw.setdefault("integrator", "fixed")
    return GreenLightVecEnv(B, **kw)


def f32_close(a, b):
    """a float32 (device), b float64 (oracle): equal after rounding b to float32, within one float32 ulp"""
    b32 = np.asarray(b, dtype=np.float64).astype(np.float32)
    return bool(np.all(np.abs(a.astype(np.float64) - b32.astype(np.float64)) <= 1.01 * np.spacing(np.abs(b32)).astype(np.float64) + 1e-30))


def state_obs_ref(seed, env_id, ctr):
    out = np.zeros(27)
    for b in range(14):
        r = philox_ref.philox4x32_10(96 + b, ctr & 0xFFFFFFFF, env_id & 0xFFFFFFFF, (env_id >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF,
                                     (seed >> 32) & 0xFFFFFFFF)
        out[2 * b] = philox_ref.u01(r[0], r[1])
        if 2 * b + 1 < 27:
            out[2 * b + 1] = philox_ref.u01(r[2], r[3])
    return out


STACKS = [
    ["IndoorClimateObservations", "BasicCropObservations", "ControlObservations", "WeatherObservations", "TimeObservations"],
    ["TimeObservations", "ControlObservations", "IndoorClimateObservations"],
    ["WeatherForecastObservations", "BasicCropObservations", "IndoorClimateObservations"],
    ["WeatherObservations", "IndoorClimateObservations", "WeatherForecastObservations", "TimeObservations"],
    ["StateObservations", "IndoorClimateObservations", "WeatherForecastObservations"],
    ["ControlObservations", "StateObservations"],
]


@pytest.mark.parametrize("si", range(len(STACKS)))
@pytest.mark.parametrize("role_warps", [1, 2])
def test_observation_module_stacks(si, role_warps, weather0, params64):
    """observations.py:9-182 / tomato_env.py:77-96: any ordered subset of the seven modules.  Rows, rewards (which read
    obs[0:3] of whatever stack is configured, rewards.py:191-198), terminal observation and in-place reset against the oracle
    with the same stack (itself pinned on the reference env's traces, test_observation_module_stacks_match_reference_env);
    StateObservations (random numbers in the reference) against the numpy restatement of the device's Philox stream."""
    mods = STACKS[si]
    B, seed, off = 3, 77, 1000
    N_short = 6  # short season so termination / auto-reset is reached
    base = dict(season_length=N_short / 96.0)  # exactly representable: N = int(season_length * 86400 / dt)
    env = make_env(B, n_sub=300, observation_modules=mods, seed=seed, env_id_offset=off, role_warps=role_warps, base_env_params=base,
                   weather_tables=weather0)
    assert env.N == N_short
    cfg = ob.default_cfg(n_sub=300, N=N_short, obs_modules=mods)
    orc = [ob.OracleEnv(weather0, params64, cfg) for _ in range(B)]
    has_state = "StateObservations" in mods
    assert env.obs_dim == orc[0].nobs == env.observation_space.shape[0] == len(env.get_obs_names())
    obs = env.reset()
    for b in range(B):
        if has_state:
            orc[b].set_state_obs(state_obs_ref(seed, off + b, 0x80000000))
        assert f32_close(obs[b], orc[b].reset()), (mods, b)
    rng = np.random.default_rng(si)
    for s in range(N_short + 3):
        a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
        obs, rew, done, infos = env.step(a)
        ctr = env.state_dict()["step_ctr"] - 1  # counter the step just used
        for b in range(B):
            if has_state:
                orc[b].set_state_obs(state_obs_ref(seed, off + b, int(ctr[b])))
            o, r, dn, _ = orc[b].step(action=a[b])
            assert bool(done[b]) == dn and abs(float(env.reward_t[b]) - r) <= 1e-9, (mods, s, b)
            if dn:
                assert f32_close(infos[b]["terminal_observation"], o), (mods, s, b)
                if has_state:
                    orc[b].set_state_obs(state_obs_ref(seed, off + b, (int(ctr[b]) + 0x80000000) & 0xFFFFFFFF))
                assert f32_close(obs[b], orc[b].reset()), (mods, s, b)
            else:
                assert f32_close(obs[b], o), (mods, s, b)
    env.close()


def test_observation_stack_errors():
    with pytest.raises(ValueError):
        make_env(2, observation_modules=["ControlObservations", "ControlObservations"])
    with pytest.raises(ValueError):
        make_env(2, observation_modules=["NoSuchObservations"])
    from glgym import _lib
    import ctypes as C
    L = _lib.load()
    cfg = _lib.GlgConfig()
    L.glg_default_config(C.byref(cfg))
    cfg.obs_modules[0], cfg.obs_modules[1] = 4, 9
    h = C.c_void_p()
    assert L.glg_create(C.byref(cfg), C.byref(h)) == _lib.GLG_ERR_ARG


# ------------------------------------------------------------------------------------------------ reference unit tests
@pytest.fixture()
def tomato():
    from glgym import TomatoEnv
    env = TomatoEnv(base_env_params=dict(season_length=60))
    env.reset(seed=42)
    yield env
    env.close()


class TestTomatoEnvReferenceSuite:
    """/root/reference/tests/env_test.py:17-92 (TestTomatoEnv), restated against glgym.TomatoEnv: same calls, same assertions."""

    def test_reward_normalisation(self, tomato):
        env = tomato
        obs, info = env.reset(seed=42)
        max_reward = 0.328 * 900 * 1e-6 / 0.065 * 1.6
        assert abs(env.reward.max_profit - max_reward) < 5e-8  # assertAlmostEqual: 7 places
        assert env.reward.variable_costs == 0
        action = np.ones(env.nu) * 1
        env.u = np.ones(env.nu) * 1
        obs, reward, terminated, truncated, info = env.step(action)
        assert np.isfinite(env.reward.scale_reward(env.reward.profit, env.reward.min_profit, env.reward.max_profit))
        violations = env.reward.output_violations()
        scaled = env.reward.scale_reward(violations, env.reward.min_state_violations, env.reward.max_state_violations)
        assert scaled.shape == (3,) and np.all(scaled >= 0)
        assert np.array_equal(env.u, np.ones(env.nu))  # u was 1, action +1 -> clipped at u_max

    def test_reset(self, tomato):
        obs, info = tomato.reset(seed=42)
        assert len(obs) == tomato.observation_space.shape[0]
        assert tomato.timestep == 0
        assert not tomato.terminated and info == {}

    def test_step(self, tomato):
        env = tomato
        env.reset()
        action = env.action_space.sample()
        obs, reward, terminated, truncated, info = env.step(action)
        assert len(obs) == env.observation_space.shape[0]
        assert isinstance(reward, (int, float))
        assert env.timestep == 1 and truncated is False
        assert set(info) >= {"EPI", "revenue", "variable_costs", "fixed_costs", "co2_cost", "heat_cost", "elec_cost", "temp_violation",
                             "co2_violation", "rh_violation", "lamp_violation", "controls"}

    def test_reward(self, tomato):
        env = tomato
        env.reset()
        action = np.ones(env.nu) * -1
        obs, reward, terminated, truncated, info = env.step(action)
        assert isinstance(reward, (int, float))
        assert env.reward.variable_costs == 0

    def test_action_scaling(self, tomato):
        env = tomato
        action = env.action_space.sample()
        scaled = env.action_to_control(action)
        assert np.all(scaled >= env.u_min) and np.all(scaled <= env.u_max)

    def test_episode_termination(self):
        from glgym import TomatoEnv
        env = TomatoEnv(base_env_params=dict(season_length=2))  # 2 days: 193 steps (the rule is N + 1)
        env.reset()
        limit = 2 * 86400 // 900 + 1
        terminated, steps = False, 0
        while not terminated and steps < limit:
            _, _, terminated, _, _ = env.step(env.action_space.sample())
            steps += 1
        assert steps == limit and terminated
        env.close()


def test_tomato_env_extras(weather0, params64):
    """step_raw_control (tomato_env.py:148-173), step_raw_control_pipeinput (:175-191: nominal parameters, no clock update),
    set_crop_state (:224-229), set_seed / eval-mode bookkeeping (base_env.py:160-170, tomato_env.py:240-244)."""
    from glgym import TomatoEnv
    env = TomatoEnv(n_sub=300, integrator="fixed", eval_options=dict(eval_days=[0], eval_years=[2009], location="Bleiswijk", data_source="GL"),
                    base_env_params=dict(training=False))
    obs, _ = env.reset(seed=3)
    assert env.eval_idx == 1 and env.location == "Bleiswijk" and env.growth_year == 2009 and env.start_day == 0.0
    env.reset()
    assert env.eval_idx == 2
    env._reset_eval_idx()
    assert env.eval_idx == 0
    orc = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=300))
    u = np.array([0.3, 0.1, 0.8, 0.05, 0.0, 0.2])
    o, r, term, trunc, info = env.step_raw_control(u)
    oo, ro, dn, io = orc.step(control=u)
    assert f32_close(o, oo) and abs(r - ro) <= 1e-9 and np.array_equal(info["controls"], u) and abs(info["EPI"] - io[0]) <= 1e-12
    env.set_crop_state(10.0, 1.0e5, 2.6e5, 6.0e4, 3.2e3)
    x = env.x
    assert list(x[22:27]) == [10.0, 1.0e5, 2.6e5, 6.0e4, 3.2e3]
    hod, doy, k = env.hour_of_day, env.day_of_year, env.timestep
    y, term = env.step_raw_control_pipeinput(u)
    yo, bad = ob.evalf(x, u, weather0[k], params64, 900.0, 300)
    assert rel_err(np.asarray(y), yo) <= 1e-9 and env.timestep == k + 1 and not term
    assert env.hour_of_day == hod and env.day_of_year == doy  # the reference does not advance the clock here
    names = env.get_obs_names()
    assert len(names) == 263 and names[:4] == ["co2_air", "temp_air", "rh_air", "pipe_temp"]
    env.close()


# ------------------------------------------------------------------------------------------------ state, guards, scale
def test_state_dict_roundtrip_and_clock(weather0, params64):
    """Checkpoint / restore of everything a step reads: two envs stepped apart, state moved over, then they stay
    bit-identical -- including the Philox stream position under parametric uncertainty; set_state(timestep=k) moves the clock;
    out-of-range timesteps / table ids are rejected."""
    from glgym._lib import GlgError
    B = 5
    rng = np.random.default_rng(5)
    ea = make_env(B, n_sub=300, uncertainty_scale=0.2, seed=9)
    eb = make_env(B, n_sub=300, uncertainty_scale=0.2, seed=9)
    ea.reset(); eb.reset()
    for s in range(7):
        ea.step(rng.uniform(-1, 1, (B, 6)).astype(np.float32))
    sd = ea.state_dict()
    assert set(sd) == {"x", "u", "timestep", "table", "time", "step_ctr", "ep_return", "ep_len", "ep_info"}
    assert list(sd["timestep"]) == [7] * B and list(sd["ep_len"]) == [7] * B and np.allclose(sd["time"][:, 1], 7 * 0.25)
    eb.load_state_dict(sd)
    for s in range(5):
        a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
        oa, ra, da, _ = ea.step(a)
        ob_, rb, db, _ = eb.step(a)
        assert np.array_equal(oa, ob_) and np.array_equal(ra, rb)
    sa, sb = ea.state_dict(), eb.state_dict()
    assert all(np.array_equal(sa[k], sb[k]) for k in sa)
    # teacher forcing with a timestep: the clock follows (ADVICE: it used to stay at the reset value)
    eb.reset()
    eb.set_state(timestep=np.full(B, 100, dtype=np.int32))
    t = eb.state_dict()["time"]
    assert np.allclose(t[:, 0], 100 * 900 / 86400.0, rtol=0, atol=1e-12) and np.allclose(t[:, 1], (100 * 0.25) % 24)
    orc = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=300))
    for s in range(100):
        orc.e.day_of_year += (900 / 86400.0) % 365
        orc.e.hour_of_day = (orc.e.hour_of_day + 0.25) % 24
    assert t[0, 0] == orc.e.day_of_year and t[0, 1] == orc.e.hour_of_day
    with pytest.raises(GlgError):
        eb.set_state(timestep=np.full(B, -1, dtype=np.int32))
    with pytest.raises(GlgError):
        eb.load_state_dict({"table": np.full(B, 3, dtype=np.int32)})
    with pytest.raises(GlgError):
        eb.load_state_dict({"timestep": np.full(B, 10**6, dtype=np.int32)})
    ea.close(); eb.close()


@pytest.mark.parametrize("role_warps", [1, 2, 3])
def test_harvest_window_guard_with_injected_states(role_warps, weather0, params64):
    """with NOMINAL parameters a state injected inside the harvest window (cLeaf / cFruit at or above their
    maxima, as set_crop_state allows for a mature crop) makes the leaf / fruit harvest stiff (smoothHar, aux_states.hpp:75-79);
    the micro-step guard must act in every kernel variant, not only under parametric uncertainty."""
    B = 40
    env = make_env(B, n_sub=600, role_warps=role_warps)
    env.reset()
    x, u, k = env.get_state()
    c_leaf_max, c_fruit_max = params64[144], params64[145]  # laiMax / sla = 1.128e5 ; 3e6 (parameters.py)
    rng = np.random.default_rng(1)
    x[:, 23] = rng.uniform(0.98, 1.10, B) * c_leaf_max
    x[:, 25] = np.where(rng.uniform(size=B) < 0.5, x[:, 25], rng.uniform(0.999, 1.01, B) * c_fruit_max)
    env.set_state(x=x)
    orc = []
    for b in range(0, B, 5):
        o = ob.OracleEnv(weather0, params64, ob.default_cfg(n_sub=600))
        o.e.x[:] = list(x[b])
        o.e.x_prev[:] = list(x[b])
        orc.append(o)
    a = rng.uniform(-1, 1, (2, B, 6)).astype(np.float32)
    micro = 0
    for s in range(2):
        env.step(a[s])
        xg, _, _ = env.get_state()
        for j, b in enumerate(range(0, B, 5)):
            orc[j].step(action=a[s, b])
            micro += orc[j].e.n_micro
            assert rel_err(xg[b], orc[j].x) <= 1e-9, (s, b, rel_err(xg[b], orc[j].x))
    assert micro > 2 * len(orc) * 600  # the guard really split substeps somewhere
    env.close()


def test_headline_kernel_matches_oracle_at_full_batch(weather0, params64):
    """BASELINE config 2 at its full size through the kernel the auto-pick really uses there (the 4096-env test
    compared two kernels with each other): 4096 envs, random actions, both integrator contracts; 48 envs spread over the batch
    (first / last CTA included) against the oracle, per step <= 1e-9."""
    B = 4096
    rng = np.random.default_rng(8)
    idx = np.unique(np.concatenate([[0, 1, 31, 32, B - 33, B - 1], rng.integers(0, B, 42)]))
    for integ, sg, n_sub in (("fixed", 0, 600), ("graded", 3, 260)):
        env = make_env(B, integrator=integ)  # role_warps auto
        env.reset()
        cfg = ob.default_cfg(n_sub=n_sub, stiff_guard=sg)
        orc = {int(b): ob.OracleEnv(weather0, params64, cfg) for b in idx}
        for s in range(2):
            a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
            obs, rew, done, _ = env.step(a)
            xg, ug, _ = env.get_state()
            for b, o in orc.items():
                oo, r, dn, _ = o.step(action=a[b])
                assert rel_err(xg[b], o.x) <= 1e-9 and abs(float(env.reward_t[b]) - r) <= 1e-9 and f32_close(obs[b], oo), (integ, s, b)
        env.close()


def test_nccl_stats_allreduce_entry_points():
    """glg_nccl_unique_id / glg_nccl_init / glg_allreduce_stats on one rank: a world of 1 must leave the statistics
    unchanged; without a communicator the call fails loudly.  (bench.py --gpus N runs it across N ranks on hardware.)"""
    from glgym._lib import GlgError
    env = make_env(64, n_sub=300, base_env_params=dict(season_length=3 / 96.0), weather_tables=np.load(os.path.join(GOLD, "weather_golden.npz"))["sample_sd0"])
    env.reset()
    with pytest.raises(GlgError):
        env.allreduce_stats()
    a = np.zeros((64, 6), dtype=np.float32)
    for s in range(4):
        env.step(a)
    before = env.episode_stats()
    assert before["episodes"] == 64
    env.init_stats_allreduce()
    env.allreduce_stats()
    torch.cuda.synchronize()
    after = env.episode_stats()
    assert after == before
    env.close()


@pytest.mark.parametrize("mods", [None, ["WeatherObservations", "WeatherForecastObservations", "IndoorClimateObservations", "TimeObservations"],
                                  ["ControlObservations", "BasicCropObservations"]])
def test_split_observation_host_path(mods, weather0):
    """step_split(): only the per-env columns cross PCIe; the forecast block is read from the host's copy of the weather bank.
    Rows assembled from the split form equal step()'s rows bit for bit -- through an episode end with in-place reset, with the
    forecast block in the middle of the row, and for a stack without one."""
    B = 37
    kw = dict(n_sub=300, observation_modules=mods, base_env_params=dict(season_length=5 / 96.0), weather_tables=weather0, seed=4)
    ea, eb = make_env(B, **kw), make_env(B, **kw)
    ea.reset(); eb.reset()
    rng = np.random.default_rng(2)
    for s in range(9):
        a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
        oa, ra, da, ia = ea.step(a)
        so, rb, db, ib = eb.step_split(a)
        assert np.array_equal(so.full(), oa) and np.array_equal(ra, rb) and np.array_equal(da, db), s
        assert so.head.shape == (B, ea.obs_dim - (240 if ea.forecast_offset >= 0 else 0))
        if ea.forecast_offset >= 0:
            assert np.array_equal(so.forecast(3), oa[3, ea.forecast_offset:ea.forecast_offset + 240])
        if da.any():
            assert all(np.array_equal(ia[i]["terminal_observation"], ib[i]["terminal_observation"]) for i in np.nonzero(da)[0])
    assert da.sum() == 0 and s == 8
    ea.close(); eb.close()


@pytest.mark.parametrize("mods", [None, ["WeatherObservations", "WeatherForecastObservations", "IndoorClimateObservations", "TimeObservations"]])
@pytest.mark.parametrize("B", [37, 20000])
def test_overlapped_host_observation_path(mods, B, weather0):
    """glg_step_host, default mode: the forecast block of every row is written by the host from its copy of the weather bank while
    the kernel runs (predicted from the previous step's timestep / table), the other columns cross PCIe, and the prediction is
    verified after the step.  The array equals the full device->host copy (host_obs="copy") and the device's own obs_t bit for
    bit -- through episode ends with in-place resets into other weather tables, after reset(), tensor steps and set_state in
    between (which invalidate the prediction), with the forecast block last or in the middle of the row, with page-locked and
    pageable caller buffers, single- and multi-threaded host loops (B = 20 000 moves 19 MB per step)."""
    tabs = np.stack([weather0, weather0[::-1].copy(), weather0 * 0.5])
    kw = dict(n_sub=300, observation_modules=mods, base_env_params=dict(season_length=5 / 96.0), weather_tables=tabs,
              table_start_days=np.array([0.0, 4.0, 9.0]), seed=11)
    ea, eb = make_env(B, host_obs="overlap", **kw), make_env(B, host_obs="copy", **kw)
    assert ea.host_obs == "overlap"
    ea.reset(); eb.reset()
    rng = np.random.default_rng(5)
    n_done = 0
    for s in range(14):
        a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
        if s == 8:  # something other than step() moves the envs: the host's prediction is stale / invalid
            at = torch.as_tensor(a, device="cuda")
            ea.step_tensor(at); eb.step_tensor(at)
            continue
        if s == 10:
            k = np.full(B, 2, dtype=np.int32); k[::3] = 4
            ea.set_state(timestep=k); eb.set_state(timestep=k)
        if s == 12:
            ea.reset(); eb.reset()
        oa, ra, da, ia = ea.step(a)
        ob_, rb, db, ib = eb.step(a)
        assert np.array_equal(oa, ob_), (s, np.argwhere(oa != ob_)[:4])
        assert np.array_equal(oa, ea.obs_t.cpu().numpy()), s
        assert np.array_equal(ra, rb) and np.array_equal(da, db)
        n_done += int(da.sum())
        if da.any():
            assert len(set(ea.table_t.cpu().numpy().tolist())) > 1  # the resets drew different weather tables
            i = int(np.nonzero(da)[0][0])
            assert np.array_equal(ia[i]["terminal_observation"], ib[i]["terminal_observation"])
    assert n_done >= B
    # the C entry with pageable caller buffers
    L = ea._lib
    obs = np.empty((B, ea.obs_dim), dtype=np.float32); rew = np.empty(B); done = np.empty(B, dtype=np.uint8)
    a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
    assert L.glg_step_host(ea._h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data) == 0
    ob_, rb, db, _ = eb.step(a)
    assert np.array_equal(obs, ob_) and np.array_equal(rew.astype(np.float32), rb) and np.array_equal(done.astype(bool), db)
    assert L.glg_set_host_obs_mode(ea._h, 7) != 0
    ea.close(); eb.close()


def test_host_step_follows_setters(weather0):
    """Setters between host steps (new Philox key, new weather bank -- which also replaces the host's float32 copy the forecast
    blocks are written from) take effect in the overlapped host path exactly as in the full-copy one: identical through 12 steps
    with a reseed and a new weather bank in the middle.  (Replaying the host step's four device operations as a CUDA graph was
    tried on top of this test and measured no faster -- 852-867 us against 846 us per step at B = 4096 -- so it is not in the build.)"""
    B = 64
    kw = dict(n_sub=8, uncertainty_scale=0.2, weather_tables=weather0, seed=3, base_env_params=dict(season_length=1.0))
    ea, eb = make_env(B, host_obs="overlap", **kw), make_env(B, host_obs="copy", **kw)
    ea.reset(); eb.reset()
    rng = np.random.default_rng(0)
    for s in range(12):
        if s == 6:
            ea.reseed(77); eb.reseed(77)   # the Philox key is a kernel argument
        if s == 9:
            W2 = np.ascontiguousarray(weather0 * 1.01)
            for e in (ea, eb):
                assert e._lib.glg_set_weather(e._h, W2.ctypes.data, 1, W2.shape[0], None) == 0
        a = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
        oa, ra, da, _ = ea.step(a)
        ob_, rb, db, _ = eb.step(a)
        assert np.array_equal(oa, ob_) and np.array_equal(ra, rb) and np.array_equal(da, db), s
    assert torch.equal(ea.state_t, eb.state_t)
    ea.close(); eb.close()
