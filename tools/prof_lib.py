"""ncu driver with an alternative build of the library: python tools/prof_lib.py <path.so> <role_warps> <B> <n_sub> [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from glgym.vec_env import GreenLightVecEnv
rw, B, n_sub = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
env = GreenLightVecEnv(B, n_sub=n_sub, integrator="fixed", role_warps=rw); env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(steps):
    env.step_tensor(torch.rand(B, 6, device="cuda", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("ok", env.launch_count())
