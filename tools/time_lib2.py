"""Step time of both integrator contracts and both layouts with an alternative build: python tools/time_lib2.py <path.so>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from glgym.vec_env import GreenLightVecEnv
tag = os.path.basename(sys.argv[1])
for B, integ in ((4096, "fixed"), (4096, "graded"), (262144, "graded")):
    env = GreenLightVecEnv(B, integrator=integ); env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(3): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if B <= 65536 else 4
    e0.record()
    for _ in range(n): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{tag} B={B} {integ}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s", flush=True)
    env.close()
