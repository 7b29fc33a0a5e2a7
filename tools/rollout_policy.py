"""BASELINE config 4 shape: PPO-style rollouts with a policy network in the loop, everything on the device.

python tools/rollout_policy.py [--envs 65536] [--steps 32] [--precision fp64|fp32] [--wrapper kernels|torch]
(under torchrun: one rank per GPU, independent env shards, env-steps/s summed over ranks)

Policy / value networks follow gl_gym/configs/agents/ppo.yml (pi 256x3, vf 512x3, SiLU, log_std_init = log 1); weights are
random (no checkpoint is available offline).  Per step: policy + value forward, sample and clip the action like SB3, env step,
VecNormalize statistics + normalisation of observations and rewards into the rollout buffer; after --steps steps the advantages
(GAE, gamma / gae_lambda of ppo.yml).  --wrapper kernels = glgym.rollout.DeviceRollout (csrc/glg_rollout.cuh), torch = the eager
torch restatement of round 1 (glgym.normalize).  Reports the step time split into policy, env step and wrapper.
The networks are plain torch modules (library GEMMs): they are the caller of the hot path, not part of it.
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
from glgym.normalize import DeviceVecNormalize, EpisodeMonitor
from glgym.rollout import DeviceRollout


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
    ap.add_argument("--integrator", default="graded")
    ap.add_argument("--wrapper", default="kernels", choices=["kernels", "torch"])
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    env = GreenLightVecEnv(a.envs, integrator=a.integrator, device=local, seed=0, env_id_offset=rank * a.envs, precision=a.precision,
                           uncertainty_scale=a.uncertainty)
    T, gamma, lam = a.steps, 0.9631, 0.9470
    torch.manual_seed(rank)
    pi, vf = mlp(env.obs_dim, [256] * 3, 6).to(dev), mlp(env.obs_dim, [512] * 3, 1).to(dev)
    log_std = torch.zeros(6, device=dev)

    @torch.no_grad()
    def act(obs):
        mean = pi(obs)
        value = vf(obs).squeeze(-1)
        action = mean + torch.exp(log_std) * torch.randn_like(mean)
        return torch.clamp(action, -1.0, 1.0), value

    ev = lambda: torch.cuda.Event(enable_timing=True)
    values = torch.zeros(T + 1, a.envs, device=dev)
    if a.wrapper == "kernels":
        roll = DeviceRollout(env, T, gamma=gamma, gae_lambda=lam)
        obs = roll.reset()
        step = lambda action: roll.step(action)[0]
    else:
        venv, mon = DeviceVecNormalize(env, gamma=gamma), EpisodeMonitor(a.envs, dev)
        obs = venv.reset_tensor()

        def step(action):
            o, r, d = venv.step_tensor(action)
            mon.update(env.reward_t, d)
            return o
    for w in range(2):  # warm-up rollout(s)
        for t in range(min(T, 4)):
            obs = step(act(obs)[0])
        if a.wrapper == "kernels":
            obs = roll.begin()
    torch.cuda.synchronize()
    marks = [(ev(), ev(), ev()) for _ in range(T)]
    e0, e1 = ev(), ev()
    e0.record()
    for t in range(T):
        marks[t][0].record()
        action, values[t] = act(obs)
        marks[t][1].record()
        if a.wrapper == "kernels":
            env.step_tensor(action)
            marks[t][2].record()
            roll._store(roll.t)
            roll.t += 1
            obs = roll.obs[roll.t]
        else:
            raw = env.step_tensor(action)
            marks[t][2].record()
            obs = venv._step(*raw)[0]
            mon.update(env.reward_t, raw[2])
    values[T] = act(obs)[1]
    if a.wrapper == "kernels":
        roll.finish(values)
    e1.record()
    torch.cuda.synchronize()
    t_pol = sum(m[0].elapsed_time(m[1]) for m in marks) / T
    t_env = sum(m[1].elapsed_time(m[2]) for m in marks) / T
    total = e0.elapsed_time(e1)
    ms = torch.tensor([total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        rate = world * a.envs * T / (ms.item() * 1e-3)
        per = ms.item() / T
        print(f"policy-in-the-loop rollout ({a.wrapper} wrapper): {world} GPU x {a.envs} envs, {a.precision}, {a.integrator}, {T} steps: "
              f"{per:.2f} ms/step = policy {t_pol:.2f} + env step {t_env:.2f} + normalise / buffer / GAE {per - t_pol - t_env:.2f}; "
              f"{rate:.3e} env-steps/s (finite obs: {bool(torch.isfinite(obs).all())})")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
