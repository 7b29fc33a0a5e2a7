"""ncu driver for the fp32 throughput mode: python tools/prof_step_fp32.py <B> <n_sub> [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
B, n_sub = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = GreenLightVecEnv(B, n_sub=n_sub, integrator="fixed", precision="fp32"); env.reset_tensor()
A = torch.rand(B, 6, device="cuda") * 2 - 1
for _ in range(steps): env.step_tensor(A)
torch.cuda.synchronize()
