"""PPO on the batched GPU env with nothing leaving the device: env step (fused CUDA kernel), VecNormalize statistics +
normalisation + rollout buffer + GAE (`glgym.rollout.DeviceRollout`, csrc/glg_rollout.cuh), policy / value networks and the
clipped-surrogate update in plain torch.  Hyper-parameters and network shapes follow the reference's
gl_gym/configs/agents/ppo.yml (pi 256x3, vf 512x3, SiLU, gamma 0.9631, gae_lambda 0.9167, clip 0.2, ent 0.05434, vf 0.8225,
max_grad_norm 0.3, lr 2e-5 -- raised here by default because this demo runs minutes, not 2e7 steps).  It replaces the
reference's SB3 pipeline `SubprocVecEnv -> VecMonitor -> VecNormalize -> PPO.collect_rollouts` (gl_gym/RL/utils.py:44-69);
with stable-baselines3 installed, examples/sb3_ppo.py keeps SB3 itself on top of the numpy VecEnv protocol instead.

    python examples/ppo_device_rollout.py [--n-envs 4096] [--n-steps 32] [--iters 30] [--lr 3e-4]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.rollout import DeviceRollout
from glgym.vec_env import GreenLightVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--n-envs", type=int, default=4096)
ap.add_argument("--n-steps", type=int, default=32)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--epochs", type=int, default=4)
ap.add_argument("--minibatch", type=int, default=16384)
ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--uncertainty-scale", type=float, default=0.0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)


def mlp(i, hidden, o):
    layers, d = [], i
    for h in hidden:
        layers += [torch.nn.Linear(d, h), torch.nn.SiLU()]
        d = h
    return torch.nn.Sequential(*layers, torch.nn.Linear(d, o))


env = GreenLightVecEnv(a.n_envs, uncertainty_scale=a.uncertainty_scale, seed=666)
T, B, GAMMA, LAM = a.n_steps, a.n_envs, 0.9631, 0.9167
roll = DeviceRollout(env, T, gamma=GAMMA, gae_lambda=LAM, clip_obs=10.0)
pi, vf = mlp(env.obs_dim, [256] * 3, 6).to(dev), mlp(env.obs_dim, [512] * 3, 1).to(dev)
log_std = torch.nn.Parameter(torch.zeros(6, device=dev))
opt = torch.optim.Adam(list(pi.parameters()) + list(vf.parameters()) + [log_std], lr=a.lr)
LOG2PI = 0.9189385332046727


def logp_of(mean, action):
    return (-0.5 * ((action - mean) / log_std.exp()) ** 2 - log_std - LOG2PI).sum(-1)


actions = torch.zeros(T, B, 6, device=dev)
logps = torch.zeros(T, B, device=dev)
values = torch.zeros(T + 1, B, device=dev)
obs = roll.reset()
t_start, env_steps = time.time(), 0
for it in range(a.iters):
    env.episode_stats(clear=True)
    raw_rew = 0.0
    with torch.no_grad():
        for t in range(T):
            mean = pi(obs)
            values[t] = vf(obs).squeeze(-1)
            act = mean + log_std.exp() * torch.randn_like(mean)
            actions[t], logps[t] = act, logp_of(mean, act)
            obs, _, _ = roll.step(act.clamp(-1.0, 1.0))   # SB3 clips to the action box when stepping, stores the raw sample
            raw_rew += env.reward_t.mean().item()
        values[T] = vf(obs).squeeze(-1)
        adv, ret = roll.finish(values)
    env_steps += T * B
    b_obs, b_act, b_logp = roll.obs[:T].reshape(T * B, -1), actions.reshape(T * B, 6), logps.reshape(T * B)
    b_adv, b_ret = adv.reshape(T * B), ret.reshape(T * B)
    for ep in range(a.epochs):
        perm = torch.randperm(T * B, device=dev)
        for lo in range(0, T * B, a.minibatch):
            idx = perm[lo:lo + a.minibatch]
            A_ = b_adv[idx]
            A_ = (A_ - A_.mean()) / (A_.std() + 1e-8)
            mean = pi(b_obs[idx])
            ratio = (logp_of(mean, b_act[idx]) - b_logp[idx]).exp()
            pg = -torch.min(ratio * A_, ratio.clamp(0.8, 1.2) * A_).mean()
            v_loss = torch.nn.functional.mse_loss(vf(b_obs[idx]).squeeze(-1), b_ret[idx])
            entropy = (log_std + 0.5 + LOG2PI).sum()
            loss = pg + 0.8225 * v_loss - 0.05434 * entropy
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(pi.parameters()) + list(vf.parameters()) + [log_std], 0.3)
            opt.step()
    obs = roll.begin()
    st = env.episode_stats()
    print(f"iter {it:3d}: mean raw reward per step {raw_rew / T:+.4f}   finished episodes {int(st['episodes'])}   "
          f"{env_steps / (time.time() - t_start):.3e} env-steps/s incl. updates", flush=True)
env.close()
