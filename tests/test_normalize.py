"""device-side VecNormalize / VecMonitor equivalents against a numpy restatement of the SB3 2.6.0 algorithms
(common/running_mean_std.py, common/vec_env/vec_normalize.py, common/vec_env/vec_monitor.py; SB3 is pinned by the
reference's requirements.txt:49 and not installed here, so the restatement below is the oracle)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from glgym.normalize import DeviceVecNormalize, EpisodeMonitor, RunningMeanStd  # noqa: E402


class NpRunningMeanStd:
    def __init__(self, shape=(), epsilon=1e-4):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), epsilon

    def update(self, arr):
        bm, bv, bc = np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class NpVecNormalize:
    def __init__(self, n, obs_dim, gamma, clip_obs=10.0, clip_reward=10.0, epsilon=1e-8):
        self.obs_rms, self.ret_rms = NpRunningMeanStd((obs_dim,)), NpRunningMeanStd(())
        self.returns, self.gamma, self.co, self.cr, self.eps = np.zeros(n), gamma, clip_obs, clip_reward, epsilon

    def norm_obs(self, obs):
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.eps), -self.co, self.co).astype(np.float32)

    def reset(self, obs):
        self.returns[:] = 0
        self.obs_rms.update(obs)
        return self.norm_obs(obs)

    def step(self, obs, rew, done):
        self.obs_rms.update(obs)
        nobs = self.norm_obs(obs)
        self.returns = self.returns * self.gamma + rew
        self.ret_rms.update(self.returns)
        nrew = np.clip(rew / np.sqrt(self.ret_rms.var + self.eps), -self.cr, self.cr)
        self.returns[done] = 0
        return nobs, nrew


class FakeEnv:
    """Tensor API of GreenLightVecEnv fed from a script (CPU)."""

    def __init__(self, obs, rew, done):
        self.script, self.k = (obs, rew, done), 0
        self.num_envs, self.obs_dim = obs.shape[1], obs.shape[2]
        self.obs_t = torch.as_tensor(obs[0])
        self.terminal_obs_t = torch.as_tensor(obs[0])

    def reset_tensor(self):
        self.k = 0
        return torch.as_tensor(self.script[0][0])

    def step_tensor(self, actions, noise=None):
        self.k += 1
        o, r, d = (torch.as_tensor(a[self.k]) for a in self.script)
        return o, r, d.to(torch.uint8)


def test_running_mean_std_matches_batch_statistics():
    rng = np.random.default_rng(0)
    data = rng.normal(3.0, 2.0, (1000, 5))
    rms = RunningMeanStd((5,))
    for chunk in np.split(data, 10):
        rms.update(torch.as_tensor(chunk))
    # with the 1e-4 pseudo-count the running moments equal the pooled moments to ~1e-7
    assert np.allclose(rms.mean.numpy(), data.mean(0), atol=1e-5) and np.allclose(rms.var.numpy(), data.var(0), rtol=1e-5)


def test_device_vec_normalize_matches_sb3_restatement():
    rng = np.random.default_rng(1)
    T, n, D = 40, 64, 263
    obs = (rng.normal(0, 1, (T + 1, n, D)) * rng.uniform(0.1, 500, D) + rng.uniform(-100, 1000, D)).astype(np.float32)
    rew = rng.normal(0.3, 0.5, (T + 1, n))
    done = rng.random((T + 1, n)) < 0.05
    env = FakeEnv(obs, rew, done)
    dv = DeviceVecNormalize(env, gamma=0.9631)
    ref = NpVecNormalize(n, D, 0.9631)
    mon = EpisodeMonitor(n, "cpu")
    ep_ret, ep_len = np.zeros(n), np.zeros(n, dtype=np.int64)
    o = dv.reset_tensor()
    assert np.allclose(o.numpy(), ref.reset(obs[0].astype(np.float64)), atol=1e-6)
    for t in range(1, T + 1):
        no, nr, d = dv.step_tensor(None)
        ro, rr = ref.step(obs[t].astype(np.float64), rew[t], done[t])
        assert np.allclose(no.numpy(), ro, atol=2e-6), t
        assert np.allclose(nr.numpy(), rr, rtol=1e-12, atol=1e-12), t
        assert np.allclose(dv.returns.numpy(), ref.returns, rtol=1e-12, atol=1e-12)
        fin_r, fin_l = mon.update(torch.as_tensor(rew[t]), torch.as_tensor(done[t]))
        ep_ret += rew[t]; ep_len += 1
        assert np.allclose(fin_r.numpy(), ep_ret[done[t]]) and np.array_equal(fin_l.numpy(), ep_len[done[t]])
        ep_ret[done[t]] = 0; ep_len[done[t]] = 0
    assert np.allclose(dv.obs_rms.var.numpy(), ref.obs_rms.var, rtol=1e-12) and abs(dv.ret_rms.var.item() - ref.ret_rms.var) < 1e-12
    # unnormalize_obs inverts normalize_obs where nothing was clipped
    z = dv.normalize_obs(torch.as_tensor(obs[3]))
    back = dv.unnormalize_obs(z).numpy()
    keep = (np.abs(z.numpy()) < 10).all(axis=1)
    assert np.allclose(back[keep], obs[3][keep], rtol=2e-5, atol=1e-3)


@pytest.mark.gpu
def test_device_vec_normalize_on_cuda_env():
    """The wrapper around the real CUDA env equals the numpy restatement applied to the raw outputs of an identical env."""
    from glgym.vec_env import GreenLightVecEnv
    B = 128
    env, raw = GreenLightVecEnv(B, n_sub=600, integrator="fixed", seed=5), GreenLightVecEnv(B, n_sub=600, integrator="fixed", seed=5)
    dv, ref = DeviceVecNormalize(env, gamma=0.9631), NpVecNormalize(B, env.obs_dim, 0.9631)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    o = dv.reset_tensor()
    ro = ref.reset(raw.reset_tensor().cpu().numpy().astype(np.float64))
    assert np.allclose(o.cpu().numpy(), ro, atol=1e-6)
    for t in range(25):
        a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
        no, nr, d = dv.step_tensor(a)
        o2, r2, d2 = raw.step_tensor(a)
        eo, er = ref.step(o2.cpu().numpy().astype(np.float64), r2.cpu().numpy(), d2.cpu().numpy().astype(bool))
        assert np.allclose(no.cpu().numpy(), eo, atol=2e-6) and np.allclose(nr.cpu().numpy(), er, rtol=1e-10, atol=1e-12), t
    env.close(); raw.close()


def np_gae(rewards, values, starts, gamma, lam):
    """RolloutBuffer.compute_returns_and_advantage of SB3 2.6.0 (common/buffers.py), float32 like the buffer's arrays:
    values [T+1, B] (row T = last_values), starts [T+1, B] (row T = dones of the last step)."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = np.zeros(rewards.shape[1], dtype=np.float32)
    for t in reversed(range(T)):
        nnt = (1.0 - starts[t + 1]).astype(np.float32)
        delta = rewards[t] + np.float32(gamma) * values[t + 1] * nnt - values[t]
        last = delta + np.float32(gamma * lam) * nnt * last
        adv[t] = last
    return adv, adv + values[:T]


@pytest.mark.gpu
def test_device_rollout_kernels_match_sb3_restatement():
    """glg_rollout_* (csrc/glg_rollout.cuh): fused running statistics + normalisation into the rollout buffer + GAE, on the real
    CUDA env, against the numpy restatement of SB3's VecNormalize / RolloutBuffer applied to the raw outputs of an identical
    env.  Short season so episode ends (returns reset, episode_starts) fall inside the rollout; two consecutive rollouts
    (begin() carries slot T over)."""
    from glgym.rollout import DeviceRollout
    from glgym.vec_env import GreenLightVecEnv
    B, T, gamma, lam = 200, 12, 0.9631, 0.9470
    kw = dict(n_sub=300, integrator="fixed", seed=5, base_env_params=dict(season_length=7 / 96.0))
    env, raw = GreenLightVecEnv(B, **kw), GreenLightVecEnv(B, **kw)
    roll = DeviceRollout(env, T, gamma=gamma, gae_lambda=lam)
    ref = NpVecNormalize(B, env.obs_dim, gamma)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    o = roll.reset()
    ro = ref.reset(raw.reset_tensor().cpu().numpy().astype(np.float64))
    assert np.allclose(o.cpu().numpy(), ro, atol=2e-6) and bool((roll.episode_starts[0] == 1).all())
    n_done = 0
    for rollout in range(2):
        exp_obs, exp_rew, exp_starts = [ro], [], [roll.episode_starts[0].cpu().numpy()]
        for t in range(T):
            a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
            no, nr, d = roll.step(a)
            o2, r2, d2 = raw.step_tensor(a)
            eo, er = ref.step(o2.cpu().numpy().astype(np.float64), r2.cpu().numpy(), d2.cpu().numpy().astype(bool))
            assert np.allclose(no.cpu().numpy(), eo, atol=2e-6) and np.allclose(nr.cpu().numpy(), er, rtol=1e-6, atol=1e-7), (rollout, t)
            assert np.array_equal(d.cpu().numpy(), d2.cpu().numpy())
            exp_obs.append(eo); exp_rew.append(er.astype(np.float32)); exp_starts.append(d2.cpu().numpy().astype(np.float32))
            n_done += int(d2.sum().item())
        assert np.allclose(roll.obs.cpu().numpy(), np.stack(exp_obs), atol=2e-6)
        assert np.allclose(roll.rewards.cpu().numpy(), np.stack(exp_rew), rtol=1e-6, atol=1e-7)
        assert np.array_equal(roll.episode_starts.cpu().numpy(), np.stack(exp_starts))
        # running statistics (float64): observation columns and the discounted return
        st = roll.stats.cpu().numpy()
        assert np.allclose(st[:-1, 0], ref.obs_rms.mean, rtol=1e-9, atol=1e-9) and np.allclose(st[:-1, 1], ref.obs_rms.var, rtol=1e-7, atol=1e-12)
        assert abs(st[-1, 0] - ref.ret_rms.mean) <= 1e-10 and abs(st[-1, 1] - ref.ret_rms.var) <= 1e-9 * ref.ret_rms.var
        assert st[0, 2] == ref.obs_rms.count
        values = torch.randn(T + 1, B, device="cuda", generator=g)
        adv, ret = roll.finish(values)
        ea, er_ = np_gae(roll.rewards.cpu().numpy(), values.cpu().numpy(), roll.episode_starts.cpu().numpy(), gamma, lam)
        assert np.allclose(adv.cpu().numpy(), ea, rtol=1e-5, atol=1e-5) and np.allclose(ret.cpu().numpy(), er_, rtol=1e-5, atol=1e-5)
        first = roll.begin()
        ro = exp_obs[-1]
        assert np.array_equal(first.cpu().numpy(), roll.obs[T].cpu().numpy()) and np.array_equal(roll.episode_starts[0].cpu().numpy(), exp_starts[-1])
    assert n_done >= 2 * B  # episodes really ended inside the rollouts
    with pytest.raises(RuntimeError):
        for t in range(T + 1):
            roll.step(torch.zeros(B, 6, device="cuda"))
    env.close(); raw.close()
