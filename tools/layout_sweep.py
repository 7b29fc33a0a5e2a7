"""Latency vs throughput layout across batch sizes (default integrator): where should the auto-pick switch?  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
for B in (2048, 4096, 4736, 6144, 8192, 9472, 12288, 16384, 32768, 65536):
    row = []
    for rw in (2, 3):
        env = GreenLightVecEnv(B, role_warps=rw); env.reset_tensor()
        A = torch.rand(B, 6, device="cuda") * 2 - 1
        for _ in range(2): env.step_tensor(A)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 6
        e0.record()
        for _ in range(n): env.step_tensor(A)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        row.append(f"{'latency' if rw == 2 else 'throughput'} {ms:7.3f} ms {B / ms * 1e3:.3e}/s")
        env.close()
    print(f"B={B:6d}: " + " | ".join(row), flush=True)
