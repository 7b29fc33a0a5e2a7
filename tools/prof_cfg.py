"""Driver for ncu captures: python tools/prof_cfg.py B=4096 integrator=graded [n_sub=..] [precision=fp32] [uncertainty_scale=0.3]
[role_warps=..] [tables=19] [steps=4] [reset_only=1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
a = dict(kv.split("=") for kv in sys.argv[1:])
B, steps = int(a.pop("B", 4096)), int(a.pop("steps", 4))
kw = dict(integrator=a.pop("integrator", "graded"))
for k in ("n_sub", "role_warps"):
    if k in a: kw[k] = int(a.pop(k))
if "precision" in a: kw["precision"] = a.pop("precision")
if "uncertainty_scale" in a: kw["uncertainty_scale"] = float(a.pop("uncertainty_scale"))
if "tables" in a: kw["base_env_params"] = dict(start_train_day=0, end_train_day=int(a.pop("tables")) - 1)
reset_only = int(a.pop("reset_only", 0))
env = GreenLightVecEnv(B, **kw)
env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(0 if reset_only else steps):
    env.step_tensor(torch.rand(B, 6, device="cuda", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("ok", env.launch_count(), kw)
