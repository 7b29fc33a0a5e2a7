"""BASELINE config 4 shape: rollouts with a policy network in the loop, everything on the device.

python tools/rollout_policy.py [--envs 65536] [--steps 32] [--precision fp64|fp32]
(under torchrun: one rank per GPU, independent env shards, env-steps/s summed over ranks)

Policy / value networks follow gl_gym/configs/agents/ppo.yml (pi 256x3, vf 512x3, SiLU, log_std_init = log 1); weights are
random (no checkpoint is available offline).  Observations and rewards pass through the device VecNormalize equivalent
(glgym/normalize.py), actions are sampled from the diagonal Gaussian and clipped to the action box like SB3 does.
The networks are plain torch modules (library GEMMs): they are the caller of the hot path, not part of it.
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
from glgym.normalize import DeviceVecNormalize, EpisodeMonitor


def mlp(i, hidden, o):
    layers, d = [], i
    for h in hidden:
        layers += [torch.nn.Linear(d, h), torch.nn.SiLU()]
        d = h
    return torch.nn.Sequential(*layers, torch.nn.Linear(d, o))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--uncertainty", type=float, default=0.0)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    env = GreenLightVecEnv(a.envs, n_sub=600, integrator="fixed", device=local, seed=0, env_id_offset=rank * a.envs, precision=a.precision,
                           uncertainty_scale=a.uncertainty)
    venv = DeviceVecNormalize(env, gamma=0.9631)
    mon = EpisodeMonitor(a.envs, dev)
    torch.manual_seed(rank)
    pi, vf = mlp(env.obs_dim, [256] * 3, 6).to(dev), mlp(env.obs_dim, [512] * 3, 1).to(dev)
    log_std = torch.zeros(6, device=dev)

    @torch.no_grad()
    def act(obs):
        mean = pi(obs)
        value = vf(obs)
        action = mean + torch.exp(log_std) * torch.randn_like(mean)
        logp = (-0.5 * ((action - mean) / torch.exp(log_std)) ** 2 - log_std - 0.9189385332046727).sum(-1)
        return torch.clamp(action, -1.0, 1.0), value, logp

    obs = venv.reset_tensor()
    for _ in range(3):
        obs, rew, done = venv.step_tensor(act(obs)[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_env = 0.0
    e0.record()
    for _ in range(a.steps):
        action, value, logp = act(obs)
        obs, rew, done = venv.step_tensor(action)
        mon.update(env.reward_t, done)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        rate = world * a.envs * a.steps / (ms.item() * 1e-3)
        print(f"policy-in-the-loop rollout: {world} GPU x {a.envs} envs, {a.precision}, {a.steps} steps: "
              f"{ms.item() / a.steps:.2f} ms/step, {rate:.3e} env-steps/s (finite obs: {bool(torch.isfinite(obs).all())})")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
