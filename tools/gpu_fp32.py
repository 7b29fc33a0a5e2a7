"""fp32 throughput mode: timing + deviation from the fp64 GPU path over a number of free-running steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = 64
e64 = GreenLightVecEnv(B, n_sub=600, precision="fp64"); e32 = GreenLightVecEnv(B, n_sub=600, integrator="fixed", precision="fp32")
e64.reset_tensor(); e32.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
worst = np.zeros(28); r_err = 0.0
for s in range(steps):
    a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
    o64, r64, d64 = e64.step_tensor(a); o32, r32, d32 = e32.step_tensor(a)
    x64 = e64.state_t.cpu().numpy(); x32 = e32.state_t.cpu().numpy()
    err = (np.abs(x32 - x64) / np.maximum(np.abs(x64), 1e-3)).max(axis=1)
    worst = np.maximum(worst, err); r_err = max(r_err, float((r32 - r64).abs().max()))
    if s in (0, 9, 99, steps - 1): print(f"step {s+1}: max rel state err {err.max():.2e} (state {err.argmax()}), max |reward diff| so far {r_err:.2e}")
print("per-state max rel err over the run:"); print(np.array2string(worst, precision=1))
for Bt, rw in ((4096, 0), (65536, 0), (262144, 0)):
    env = GreenLightVecEnv(Bt, n_sub=600, integrator="fixed", precision="fp32", role_warps=rw); env.reset_tensor()
    A = torch.rand(Bt, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 4
    print(f"fp32 B={Bt}: {ms:.3f} ms/step  {Bt / ms * 1e3:.3e} env-steps/s"); env.close()
