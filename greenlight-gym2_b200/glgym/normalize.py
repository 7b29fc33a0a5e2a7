"""Device-resident equivalents of the SB3 wrappers the reference puts around its env.

The reference builds `SubprocVecEnv -> VecMonitor -> VecNormalize` (gl_gym/RL/utils.py:60-67, settings
`norm_obs / norm_reward / clip_obs / clip_reward / gamma` from gl_gym/RL/experiment_manager.py:142-147).  With the numpy
`GreenLightVecEnv.step` those SB3 wrappers keep working unchanged; this module is the tensor fast path:
the same arithmetic on CUDA tensors, so a rollout never leaves the device.

Algorithms restated from stable-baselines3 2.6.0 (pinned in the reference's requirements.txt:49, not installed here):
  * `RunningMeanStd` -- common/running_mean_std.py: parallel (Chan et al.) update of mean / population variance from
    batch moments, count initialised to epsilon = 1e-4, float64.
  * `DeviceVecNormalize` -- common/vec_env/vec_normalize.py: `step_wait` updates obs_rms with the raw observations,
    normalises and clips them, updates the discounted return `ret = ret * gamma + reward`, feeds ret_rms, divides the
    reward by sqrt(ret_rms.var + epsilon) and clips it, zeroes the return of finished envs; `reset` zeroes the returns
    and (when training) feeds the first observations to obs_rms.
  * `EpisodeMonitor` -- common/vec_env/vec_monitor.py: per-env episode return / length, reported when done.
These are torch ops (plumbing around the hot path), not hand-written kernels; the env step itself is the CUDA path.
"""
import torch


class RunningMeanStd:
    def __init__(self, shape=(), epsilon=1e-4, device="cpu"):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = float(epsilon)

    def update(self, batch):
        batch = batch.to(torch.float64)
        self.update_from_moments(batch.mean(dim=0), batch.var(dim=0, unbiased=False), batch.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + delta.square() * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class DeviceVecNormalize:
    """VecNormalize semantics over `GreenLightVecEnv`'s tensor API.  Observations come back float32 (like the wrapped
    env), rewards float64; statistics are float64."""

    def __init__(self, env, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8):
        self.env, self.training, self.norm_obs, self.norm_reward = env, training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = float(clip_obs), float(clip_reward), float(gamma), float(epsilon)
        dev = env.obs_t.device
        self.obs_rms = RunningMeanStd((env.obs_dim,), device=dev)
        self.ret_rms = RunningMeanStd((), device=dev)
        self.returns = torch.zeros(env.num_envs, dtype=torch.float64, device=dev)
        self.old_obs = self.old_reward = None

    def normalize_obs(self, obs):
        if not self.norm_obs:
            return obs
        z = (obs.to(torch.float64) - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.epsilon)
        return torch.clamp(z, -self.clip_obs, self.clip_obs).to(torch.float32)

    def unnormalize_obs(self, obs):
        if not self.norm_obs:
            return obs
        return (obs.to(torch.float64) * torch.sqrt(self.obs_rms.var + self.epsilon) + self.obs_rms.mean).to(torch.float32)

    def normalize_reward(self, reward):
        if not self.norm_reward:
            return reward
        return torch.clamp(reward / torch.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)

    def reset_tensor(self):
        obs = self.env.reset_tensor()
        self.old_obs = obs
        self.returns.zero_()
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def _step(self, obs, reward, done):
        self.old_obs, self.old_reward = obs, reward
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        nobs = self.normalize_obs(obs)
        if self.training:
            self.returns = self.returns * self.gamma + reward
            self.ret_rms.update(self.returns)
        nrew = self.normalize_reward(reward)
        self.returns = torch.where(done.bool(), torch.zeros_like(self.returns), self.returns)
        return nobs, nrew, done

    def step_tensor(self, actions, noise=None):
        return self._step(*self.env.step_tensor(actions, noise))

    def step_rule_based_tensor(self, noise=None):
        return self._step(*self.env.step_rule_based_tensor(noise))

    def terminal_obs_tensor(self):
        """Normalised `info["terminal_observation"]` rows (valid where done)."""
        return self.normalize_obs(self.env.terminal_obs_t)


class EpisodeMonitor:
    """VecMonitor bookkeeping on the device: call `update(reward, done)` after every step; returns the (returns, lengths)
    of the episodes that just finished."""

    def __init__(self, num_envs, device):
        self.ret = torch.zeros(num_envs, dtype=torch.float64, device=device)
        self.len = torch.zeros(num_envs, dtype=torch.int64, device=device)

    def update(self, reward, done):
        self.ret += reward
        self.len += 1
        d = done.bool()
        out = self.ret[d].clone(), self.len[d].clone()
        self.ret = torch.where(d, torch.zeros_like(self.ret), self.ret)
        self.len = torch.where(d, torch.zeros_like(self.len), self.len)
        return out
