"""Timing only: python tools/gpu_time.py <role_warps> [B ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
rw = int(sys.argv[1]); rl = int(os.environ.get("ROLE_LANES", "0")); Bs = [int(b) for b in sys.argv[2:]] or [4096, 65536, 262144]
for B in Bs:
    env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", role_warps=rw, role_lanes=rl); env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 4; e0.record()
    for _ in range(n): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / n
    print(f"role_warps={rw} lanes={rl} B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s  frac_of_34.2TF={B / ms * 1e3 * 2.381e6 / 34.2e12:.3f}", flush=True)
    env.close()
