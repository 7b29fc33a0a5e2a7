"""Soak: many envs, parametric uncertainty 0.3, random-walk controls, rule-based controller; graded vs fixed integrator.
Reports non-finite terminations and the spread between the two integrators. Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
B, T = 65536, 400
for label, rule in (("random-walk controls", False), ("rule-based controller", True)):
    envs = {k: GreenLightVecEnv(B, integrator=k, uncertainty_scale=0.3, seed=11) for k in ("fixed", "graded")}
    for e in envs.values(): e.reset_tensor(); e.episode_stats(clear=True)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for s in range(T):
        a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
        for e in envs.values():
            e.step_rule_based_tensor() if rule else e.step_tensor(a)
    xf, xg = envs["fixed"].state_t, envs["graded"].state_t
    fin = {k: bool(torch.isfinite(e.state_t).all()) for k, e in envs.items()}
    bad = {k: e.stats_t[14].item() for k, e in envs.items()}
    rel = ((xf - xg).abs() / xf.abs().clamp_min(1e-3)).amax(dim=0)
    micro = envs["graded"].stats_t[15].item() / (B * T)
    print(f"{label}: B={B}, {T} steps, uncertainty 0.3: finite {fin}, non-finite terminations {bad}, RK4 steps/interval graded {micro:.1f}; "
          f"per-env max rel |fixed - graded|: median {rel.median().item():.2e}, p99 {rel.quantile(0.99).item():.2e}, max {rel.max().item():.2e}", flush=True)
    for e in envs.values(): e.close()
