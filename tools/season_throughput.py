"""Full-season episodes (BASELINE configs[4] wording): env-steps/s over a whole 5761-step episode incl. the in-place reset at its
end, per 12-day block, so the seasonal variation of the step cost shows -- the micro-step guards (harvest window, transient
stiffness) make a step dearer where they trigger, and a CTA executes the largest micro-step count of its 32 envs.
    python tools/season_throughput.py [config ...]      config = c2 | sat64 | c3     (run under gpurun)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv

CONFIGS = {
    "c2": (4096, dict(), "BASELINE configs[1]: 4096 envs, fp64, nominal parameters, one weather table"),
    "sat64": (262144, dict(), "262 144 envs, fp64, nominal parameters, one weather table"),
    "c3": (262144, dict(precision="fp32", uncertainty_scale=0.3, base_env_params=dict(start_train_day=0, end_train_day=18)),
           "BASELINE configs[2]: 262 144 envs, fp32 units, uncertainty 0.3, 19 start days"),
}
for name in (sys.argv[1:] or ["c2", "c3"]):
    B, kw, desc = CONFIGS[name]
    env = GreenLightVecEnv(B, seed=0, **kw)
    env.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    A = torch.rand(16, B, 6, device="cuda", generator=g) * 2 - 1
    N, blk = env.N + 1, 1152
    env.episode_stats(clear=True)
    rows, t_all, micro_prev = [], 0.0, 0.0
    for b0 in range(0, N, blk):
        n = min(blk, N - b0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(n):
            env.step_tensor(A[(b0 + s) % 16])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        micro = env.stats_t[15].item()
        rows.append((b0, n, ms / n, B * n / ms * 1e3, (micro - micro_prev) / (B * n)))
        micro_prev, t_all = micro, t_all + ms
    st = env.episode_stats()
    print(f"{name}: {desc}; default contract, U(-1,1) actions, {N} steps = one full episode + in-place reset")
    for b0, n, msps, rate, mic in rows:
        print(f"   steps {b0:5d}..{b0 + n - 1:5d}: {msps:8.3f} ms/step  {rate:.3e} env-steps/s  RK4 steps per env-step {mic:6.1f}")
    print(f"   whole episode: {B * N / t_all * 1e3:.3e} env-steps/s ({t_all / 1e3:.1f} s); finished episodes {st['episodes']:.0f}, "
          f"non-finite terminations {env.stats_t[14].item():.0f}, state finite {bool(torch.isfinite(env.state_t).all())}", flush=True)
    env.close()
