"""Throughput with parametric uncertainty on (Philox noise + harvest guard). Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
for prec in ("fp64", "fp32"):
    for B in (4096, 65536):
        for unc in (0.0, 0.1, 0.3):
            env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", uncertainty_scale=unc, precision=prec); env.reset_tensor()
            A = torch.rand(B, 6, device="cuda") * 2 - 1
            for _ in range(3): env.step_tensor(A)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 8; e0.record()
            for _ in range(n): env.step_tensor(A)
            e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / n
            bad = int((~torch.isfinite(env.state_tensor())).any(dim=-1).sum()) if hasattr(env, "state_tensor") else -1
            print(f"{prec} B={B} unc={unc}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s  nonfinite envs {bad}", flush=True)
            env.close()
