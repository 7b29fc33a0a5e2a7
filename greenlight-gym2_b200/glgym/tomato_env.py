"""TomatoEnv -- the reference's single-environment gymnasium API (gl_gym/environments/tomato_env.py, base_env.py) as a
facade over a one-env GreenLightVecEnv, so that code and tests written against the reference's `TomatoEnv` run unchanged
on the CUDA path:

    env = TomatoEnv(reward_function, observation_modules, constraints, eval_options, reward_params, base_env_params,
                    uncertainty_scale)
    obs, info = env.reset(seed=42)
    obs, reward, terminated, truncated, info = env.step(action)          # tomato_env.py:115-146
    ... = env.step_raw_control(u)                                        # :148-173
    x, terminated = env.step_raw_control_pipeinput(u)                    # :175-191
    env.set_crop_state(cBuf, cLeaf, cStem, cFruit, tCanSum)              # :224-229
    env.set_seed(seed); env.get_obs_names(); env.action_to_control(a)

Attributes the reference exposes and its tests / scripts read: x, u, p, timestep, terminated, N, Np, nu, nx, dt, u_min, u_max,
weather_data, day_of_year, hour_of_day, obs, growth_year, start_day, eval_idx, reward.{max_profit, min_profit, profit, gains,
variable_costs, heat_costs, co2_costs, elec_costs, *_violation, scale_reward, output_violations}.

One env per object means one CUDA launch per step for a single greenhouse: this class exists for API parity (unit tests,
evaluation scripts, debugging), not for throughput -- batch with GreenLightVecEnv for that.

Differences, all documented in DESIGN.md: the year / start-day choice of reset() and the parametric noise use the library's
Philox streams keyed by the seed instead of numpy's PCG64 (same distributions, different draws); observations are float32
(the dtype of the reference's observation space) instead of float64 arrays.
"""
import numpy as np
import torch

from .model import GreenLight
from .vec_env import (DEFAULT_BASE_ENV_PARAMS, DEFAULT_CONSTRAINTS, DEFAULT_OBSERVATION_MODULES, DEFAULT_REWARD_PARAMS, INFO_KEYS,
                      GreenLightVecEnv)


class _RewardView:
    """What the reference's GreenhouseReward object exposes (rewards.py:46-231), backed by the device-computed info of the
    last step."""

    def __init__(self, env):
        self.env = env
        rp, p, dt = env.reward_params, env.p.astype(np.float64), float(env.dt)
        self.elec_price, self.heating_price, self.co2_price = rp["elec_price"], rp["heating_price"], rp["co2_price"]
        self.fruit_price, self.dmfm = rp["fruit_price"], rp["dmfm"]
        self.pen_weights, self.pen_lamp = np.array(rp["pen_weights"]), rp["pen_lamp"]
        # rewards.py:96-124: profit bounds used for min-max scaling
        self.max_profit = p[154] * dt * 1e-6 / self.dmfm * self.fruit_price
        max_heat = p[108] / p[46] * dt / 3600 * 1e-3 * self.heating_price
        max_elec = p[172] * dt / 3600 * 1e-3 * self.elec_price
        max_co2 = p[109] / p[46] * dt * 1e-6 * self.co2_price
        self.min_profit = -(max_heat + max_elec + max_co2)
        self.min_state_violations = np.zeros(3)
        self.max_state_violations = np.array([2500.0, 15.0, 15.0])  # rewards.py:90-94
        yearly = rp["fixed_greenhouse_cost"] + rp["fixed_co2_cost"] + rp["fixed_lamp_cost"] * 116 + rp["fixed_screen_cost"]
        self.fixed_costs = yearly / 365 / (86400 // env.dt)
        self._zero()

    def _zero(self):
        self.profit = self.gains = self.variable_costs = 0.0
        self.heat_costs = self.co2_costs = self.elec_costs = 0.0
        self.temp_violation = self.co2_violation = self.rh_violation = self.lamp_violation = 0.0

    def _from_info(self, info):
        self.profit, self.gains, self.variable_costs = info["EPI"], info["revenue"], info["variable_costs"]
        self.co2_costs, self.heat_costs, self.elec_costs = info["co2_cost"], info["heat_cost"], info["elec_cost"]
        self.temp_violation, self.co2_violation, self.rh_violation = info["temp_violation"], info["co2_violation"], info["rh_violation"]
        self.lamp_violation = info["lamp_violation"]

    @staticmethod
    def scale_reward(r, min_r, max_r):  # rewards.py:126-128
        return (r - min_r) / (max_r - min_r)

    def output_violations(self):  # rewards.py:185-198 on the current observation
        o = np.asarray(self.env.obs[:3], dtype=np.float64)
        lo = np.maximum(self.env.constraints_low - o, 0.0)
        hi = np.maximum(o - self.env.constraints_high, 0.0)
        return lo + hi


class TomatoEnv:
    metadata = {"render_modes": []}

    def __init__(self, reward_function="GreenhouseReward", observation_modules=None, constraints=None, eval_options=None,
                 reward_params=None, base_env_params=None, uncertainty_scale=0.0, device=0, seed=0, **vec_kwargs):
        bp = dict(DEFAULT_BASE_ENV_PARAMS)
        bp.update(base_env_params or {})
        self._venv = GreenLightVecEnv(1, reward_function=reward_function,
                                      observation_modules=list(observation_modules or DEFAULT_OBSERVATION_MODULES),
                                      constraints=dict(DEFAULT_CONSTRAINTS, **(constraints or {})), eval_options=eval_options,
                                      reward_params=dict(DEFAULT_REWARD_PARAMS, **(reward_params or {})), base_env_params=bp,
                                      uncertainty_scale=uncertainty_scale, device=device, seed=seed, auto_reset=False,
                                      info_mode="full", obs_ring=0, **vec_kwargs)
        v = self._venv
        for name in ("nx", "nu", "nd", "num_params", "dt", "c", "u_min", "u_max", "delta_u_max", "Np", "N", "season_length",
                     "location", "data_source", "weather_data_dir", "training", "train_years", "train_days", "eval_options",
                     "uncertainty_scale", "observation_space", "action_space", "observation_modules", "constraints_low",
                     "constraints_high", "reward_params", "p"):
            setattr(self, name, getattr(v, name))
        if not self.training and eval_options is not None:  # tomato_env.py:241-243
            self.location, self.data_source = eval_options["location"], eval_options["data_source"]
        self.reward = _RewardView(self)
        self.eval_idx = 0
        self.terminated = False
        self.obs = None
        self._seed = int(seed)
        self._model = None

    # ---- reference attributes that live on the device
    @property
    def x(self):
        return self._venv.get_state()[0][0]

    @x.setter
    def x(self, value):
        self._venv.set_state(x=np.asarray(value, dtype=np.float64).reshape(1, -1))

    @property
    def u(self):
        return self._venv.get_state()[1][0]

    @u.setter
    def u(self, value):
        self._venv.set_state(u=np.asarray(value, dtype=np.float64).reshape(1, -1))

    @property
    def timestep(self):
        return int(self._venv.get_state()[2][0])

    @property
    def day_of_year(self):
        return float(self._venv.time_t[0, 0].item())

    @property
    def hour_of_day(self):
        return float(self._venv.time_t[1, 0].item())

    @property
    def weather_data(self):
        return self._venv.weather_tables[int(self._venv.table_t[0].item())]

    @property
    def growth_year(self):
        return self._venv.get_attr("growth_year")[0]

    @property
    def start_day(self):
        return self._venv.get_attr("start_day")[0]

    # ---- base_env.py:160-170
    def _reset_eval_idx(self):
        self.eval_idx = 0

    def increase_eval_idx(self):
        self.eval_idx += 1

    def set_seed(self, seed):
        """Re-keys the env's random streams (year / start-day choice, parametric noise, StateObservations)."""
        self._seed = int(seed)
        self._venv.reseed(self._seed)

    # ---- tomato_env.py
    def get_obs_names(self):
        return self._venv.get_obs_names()

    def action_to_control(self, action):  # :109-113
        return np.clip(self.u + np.asarray(action) * self.delta_u_max, self.u_min, self.u_max)

    def set_crop_state(self, cBuf, cLeaf, cStem, cFruit, tCanSum):  # :224-229
        x = self.x.copy()
        x[22:27] = (cBuf, cLeaf, cStem, cFruit, tCanSum)
        self.x = x

    def reset(self, seed=None, options=None):
        """tomato_env.py:231-270: (re)seed if a seed is given, pick year and start day from the train lists (or the eval lists
        when training=False, advancing eval_idx), load that weather table, x = init_state, u = 0, timestep = 0."""
        if seed is not None:
            self.set_seed(seed)
        if not self.training:
            self.increase_eval_idx()
        obs = self._venv.reset()
        self.obs = obs[0]
        self.terminated = False
        self.reward._zero()
        return self.obs, {}

    def _finish(self, out):
        obs, rew, done, infos = out
        self.obs = obs[0]
        if done[0]:
            self.terminated = True
        info = dict(infos[0])
        info.pop("TimeLimit.truncated", None)
        info.pop("terminal_observation", None)
        self.reward._from_info(info)
        # the reward as the reference returns it: a Python float (float64); the VecEnv protocol's array is float32
        return self.obs, float(self._venv.reward_t[0].item()), self.terminated, False, info

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, self.nu)
        return self._finish(self._venv.step(a))

    def step_raw_control(self, control):
        return self._finish(self._venv.step_raw_control(np.asarray(control, dtype=np.float64).reshape(1, self.nu)))

    def step_raw_control_pipeinput(self, control):
        """tomato_env.py:175-191: one integration with the NOMINAL parameters and nothing else -- no noise, no clock update,
        no observation; returns (x, terminated)."""
        if self._model is None:
            self._model = GreenLight(self.nx, self.nu, self.nd, self.num_params, self.dt, n_sub=self._venv.n_sub,
                                     integrator=self._venv.integrator, device=self._venv.device_index)
        st = self._venv.state_dict()
        k = int(st["timestep"][0])
        u = np.asarray(control, dtype=np.float64).reshape(self.nu)
        x_next = np.asarray(self._model.evalF(st["x"][0], u, self.weather_data[k], self.p.astype(np.float64)), dtype=np.float64)
        if k >= self.N:
            self.terminated = True
        self._venv.load_state_dict({"x": x_next.reshape(1, -1), "u": u.reshape(1, -1), "timestep": np.array([min(k + 1, self.N)], dtype=np.int32)})
        return x_next, self.terminated

    def close(self):
        self._venv.close()
