"""Per-phase cycle counters of kernel B (needs tools/ubench/libglgym_prof.so built with -DGLG_PROFILE_GROUPS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "ubench", "libglgym_prof.so")
import torch
from glgym.vec_env import GreenLightVecEnv
B = int(sys.argv[1]); rw = int(sys.argv[2]); prec = sys.argv[3] if len(sys.argv) > 3 else "fp64"
env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", role_warps=rw, precision=prec); env.reset_tensor()
A = torch.rand(B, 6, device="cuda") * 2 - 1
for _ in range(2): env.step_tensor(A)
torch.cuda.synchronize()
