"""Graded integrator vs the fixed 600-substep contract: throughput and executed micro-steps. Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
for B in (4096, 65536, 262144):
    for prec in ("fp64", "fp32"):
        for integ in ("fixed", "graded"):
            env = GreenLightVecEnv(B, integrator=integ, precision=prec); env.reset_tensor()
            A = torch.rand(B, 6, device="cuda") * 2 - 1
            for _ in range(2): env.step_tensor(A)
            env.episode_stats(clear=True); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 6 if B <= 65536 else 3
            e0.record()
            for _ in range(n): env.step_tensor(A)
            e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / n
            micro = env.stats_t[15].item() / (B * n) if integ == "graded" else env.n_sub
            print(f"B={B} {prec} {integ:6s}: {ms:8.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s  micro-steps/env-step {micro:.1f}", flush=True)
            env.close()
