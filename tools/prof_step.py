"""Tiny driver for ncu: python tools/prof_step.py <B> <n_sub> [steps] [role_warps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
B, n_sub = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw = {}
if len(sys.argv) > 4: kw["role_warps"] = int(sys.argv[4])
env = GreenLightVecEnv(B, n_sub=n_sub, integrator="fixed", **kw)
env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(steps):
    A = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
    env.step_tensor(A)
torch.cuda.synchronize()
print("ok", env.launch_count())
