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
