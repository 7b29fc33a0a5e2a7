"""Full free-running season: fp32 + graded and fp64 + graded against the fp64 fixed-step parity mode (64 envs, random-walk controls)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
B, T = 64, 5760
envs = {"fp64 fixed": GreenLightVecEnv(B), "fp64 graded": GreenLightVecEnv(B, integrator="graded"),
        "fp32 fixed": GreenLightVecEnv(B, precision="fp32"), "fp32 graded": GreenLightVecEnv(B, integrator="graded", precision="fp32")}
for e in envs.values(): e.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(2)
ret = {k: torch.zeros(B, dtype=torch.float64, device="cuda") for k in envs}
for s in range(T):
    a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
    for k, e in envs.items():
        o, r, d = e.step_tensor(a); ret[k] += r
ref = envs["fp64 fixed"].state_t
for k, e in envs.items():
    if k == "fp64 fixed": continue
    rel = ((e.state_t - ref).abs() / ref.abs().clamp_min(1e-3)).amax(dim=1)
    dr = (ret[k] - ret["fp64 fixed"]).abs().max().item()
    print(f"{k:12s} vs fp64 fixed after {T} steps: max rel state err {rel.max().item():.2e} (worst state {int(rel.argmax())}), "
          f"episode-return diff {dr:.2e} (return ~{ret['fp64 fixed'].mean().item():.1f})", flush=True)
