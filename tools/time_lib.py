"""Step time with an alternative build of the library: python tools/time_lib.py <path.so> <role_warps> [B ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from glgym.vec_env import GreenLightVecEnv
rw = int(sys.argv[2]); Bs = [int(b) for b in sys.argv[3:]] or [4096]
for B in Bs:
    env = GreenLightVecEnv(B, n_sub=600, role_warps=rw); env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(sys.argv[1])} role_warps={rw} B={B}: {e0.elapsed_time(e1) / 6:.3f} ms/step", flush=True)
    env.close()
