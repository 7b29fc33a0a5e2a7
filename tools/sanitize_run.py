"""Small runs of every kernel-B loop variant for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
for kw in (dict(), dict(role_warps=3), dict(integrator="graded", n_sub=6), dict(uncertainty_scale=0.3), dict(precision="fp32"),
           dict(role_warps=1)):
    kw.setdefault("n_sub", 8)
    env = GreenLightVecEnv(70, **kw); env.reset_tensor()
    A = torch.rand(70, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    env.step_rule_based_tensor() if kw.get("role_warps") != 1 else None
    torch.cuda.synchronize()
    print("ok", kw, bool(torch.isfinite(env.state_t).all()), flush=True)
    env.close()
