"""Evaluation outputs in the reference's on-disk layout.

`experiments/evaluate_baseline.py:12-37` and `experiments/evaluate_rl.py:37-72` step ONE env through a season and collect,
per step, obs[:23], the reward and eight info entries; `common/results.py` appends an `episode` column and writes CSV
(columns: the first 23 observation names, Rewards, EPI, Revenue, Heat costs, CO2 costs, Elec costs, temp_violation,
co2_violation, rh_violation, episode) -- the format `visualisations/*` reads.  The reference's stochastic evaluation
repeats that 30 times per uncertainty scale (`experiments/eval_baseline.sh`); here the repetitions are the envs of one
batch, each with its own Philox noise stream, recorded on the device and copied to the host once at the end.
"""
import numpy as np
import torch

from .vec_env import INFO_KEYS

RESULT_INFO = ("EPI", "revenue", "heat_cost", "co2_cost", "elec_cost", "temp_violation", "co2_violation", "rh_violation")
RESULT_COLUMNS_TAIL = ["Rewards", "EPI", "Revenue", "Heat costs", "CO2 costs", "Elec costs", "temp_violation", "co2_violation",
                       "rh_violation"]


def result_columns(env):
    """Column names of the reference's Results frame (evaluate_baseline.py:63-67)."""
    return list(env.get_obs_names()[:23]) + RESULT_COLUMNS_TAIL + ["episode"]


def _record(env, rec, t):
    done = env.done_t.bool()
    obs = torch.where(done[:, None], env.terminal_obs_t[:, :23], env.obs_t[:, :23])  # auto-reset: the last obs of a finished
    rec[t, :, :23] = obs                                                            # episode is the terminal observation
    rec[t, :, 23] = env.reward_t
    info = env.info_t
    for j, key in enumerate(RESULT_INFO):
        rec[t, :, 24 + j] = info[INFO_KEYS.index(key)]


def evaluate_rule_based(env, controller=None, n_steps=None):
    """One season of every env under the device rule-based controller.  Returns float64 [num_envs, n_steps, 32] on the
    host: obs[:23], reward, EPI, revenue, heat / co2 / elec costs, temp / co2 / rh violation per step."""
    n_steps = env.N + 1 if n_steps is None else int(n_steps)
    env.set_rule_controller(controller)
    env.reset_tensor()
    rec = torch.empty((n_steps, env.num_envs, 32), dtype=torch.float64, device=env.device)
    for t in range(n_steps):
        env.step_rule_based_tensor()
        _record(env, rec, t)
    return rec.permute(1, 0, 2).cpu().numpy()


def evaluate_policy(env, policy, n_steps=None, normalizer=None):
    """Same recording with actions from `policy(obs) -> actions` (CUDA tensors; deterministic actions like
    evaluate_rl.py:53-58).  `normalizer`: a `DeviceVecNormalize` in evaluation mode wrapping `env`, or None."""
    n_steps = env.N if n_steps is None else int(n_steps)
    src = normalizer if normalizer is not None else env
    obs = src.reset_tensor()
    rec = torch.empty((n_steps, env.num_envs, 32), dtype=torch.float64, device=env.device)
    for t in range(n_steps):
        with torch.no_grad():
            actions = torch.clamp(policy(obs), -1.0, 1.0)
        obs = src.step_tensor(actions)[0]
        _record(env, rec, t)
    return rec.permute(1, 0, 2).cpu().numpy()


def to_results_frame(data, columns):
    """Stack [episodes, steps, 32] into the reference's Results frame: one row per step, `episode` = episode index."""
    import pandas as pd
    data = np.asarray(data)
    e, t, c = data.shape
    flat = np.concatenate([data.reshape(e * t, c), np.repeat(np.arange(e, dtype=np.float64), t)[:, None]], axis=1)
    return pd.DataFrame(flat, columns=columns)


def save_results(data, columns, filename):
    to_results_frame(data, columns).to_csv(filename, index=False)
