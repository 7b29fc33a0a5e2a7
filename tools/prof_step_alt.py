"""prof_step.py against tools/ubench/libglgym_prof.so (an experimental build): python tools/prof_step_alt.py <B> <n_sub> [steps] [role_warps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "ubench", "libglgym_prof.so")
import torch
from glgym.vec_env import GreenLightVecEnv
B, n_sub = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw = {}
if len(sys.argv) > 4: kw["role_warps"] = int(sys.argv[4])
env = GreenLightVecEnv(B, n_sub=n_sub, integrator="fixed", **kw); env.reset_tensor()
A = torch.rand(B, 6, device="cuda") * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env.step_tensor(A); torch.cuda.synchronize(); e0.record()
for _ in range(steps): env.step_tensor(A)
e1.record(); torch.cuda.synchronize()
print(f"B={B} n_sub={n_sub}: {e0.elapsed_time(e1) / steps:.3f} ms/step")
