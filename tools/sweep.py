"""BASELINE config 5 shape: throughput sweep over the batch size, both precisions, default integrator contract (or `fixed`);
algorithmic flop from the RK4 steps the kernel reports having executed.  python tools/sweep.py [max_log2_B] [graded|fixed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import torch
from glgym import _lib
from glgym.vec_env import GreenLightVecEnv
import ctypes as C
top = int(sys.argv[1]) if len(sys.argv) > 1 else 20
integ = sys.argv[2] if len(sys.argv) > 2 else "graded"
pk64, pk32 = C.c_double(), C.c_double()
_lib.load().glg_measure_fp64_peak(0, C.byref(pk64)); _lib.load().glg_measure_fp32_peak(0, C.byref(pk32))
print(f"measured peaks: FP64 {pk64.value / 1e12:.1f} TFLOP/s, FP32 {pk32.value / 1e12:.1f} TFLOP/s; integrator {integ}; F_step = 3968 x executed RK4 steps + 529")
print(f"{'envs':>9} {'prec':>5} {'kernel':>8} {'ms/step':>9} {'env-steps/s':>12} {'frac of peak':>13}")
for lb in range(10, top + 1, 2):
    B = 1 << lb
    for prec, pk in (("fp64", pk64.value), ("fp32", pk32.value)):
        env = GreenLightVecEnv(B, integrator=integ, precision=prec); env.reset_tensor()
        A = torch.rand(B, 6, device="cuda") * 2 - 1
        n = 6 if B <= 65536 else 2
        for _ in range(2): env.step_tensor(A)
        torch.cuda.synchronize()
        env.episode_stats(clear=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): env.step_tensor(A)
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / n
        rate = B / ms * 1e3
        flop = 3968.0 * env.stats_t[15].item() / (B * n) + 529.0
        kern = "latency" if B <= 2 * 148 * 32 else "tput"
        print(f"{B:>9} {prec:>5} {kern:>8} {ms:>9.3f} {rate:>12.3e} {rate * flop / pk:>13.3f}", flush=True)
        env.close()
