"""Step time of kernel B with only some flux groups enabled (tools/ubench/libglgym_prof.so built with -DGLG_PROFILE_MASK)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "ubench", "libglgym_prof.so")
import torch
from glgym.vec_env import GreenLightVecEnv
B = int(sys.argv[1]); rw = int(sys.argv[2])
env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", role_warps=rw); env.reset_tensor()
A = torch.rand(B, 6, device="cuda") * 2 - 1
for _ in range(2): env.step_tensor(A)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): env.step_tensor(A)
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 4
print(f"mask {os.environ.get('GLG_PROF_MASK')} B={B} warps={rw}: {ms:.3f} ms/step = {ms * 1e-3 / 2400 * 1.965e9:.0f} cycles/eval")
