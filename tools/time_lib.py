"""Step time (and a kernel A cross-check) with an alternative build of the library:
python tools/time_lib.py <path.so> <role_warps> [B ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from glgym.vec_env import GreenLightVecEnv
rw = int(sys.argv[2]); Bs = [int(b) for b in sys.argv[3:]] or [4096]
tag = os.path.basename(sys.argv[1])
# cross-check against kernel A (one thread per env, same library): 3 steps, 96 envs, random actions
ea, eb = GreenLightVecEnv(96, n_sub=600, role_warps=1), GreenLightVecEnv(96, n_sub=600, integrator="fixed", role_warps=rw)
ea.reset_tensor(); eb.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(3)
worst = 0.0
for _ in range(3):
    A = torch.rand(96, 6, device="cuda", generator=g) * 2 - 1
    ea.step_tensor(A); eb.step_tensor(A)
    xa, xb = ea.state_t, eb.state_t
    worst = max(worst, float(((xa - xb).abs() / xa.abs().clamp_min(1e-3)).max()))
print(f"{tag} role_warps={rw}: max rel diff vs kernel A after 3 steps {worst:.2e}", flush=True)
ea.close(); eb.close()
for B in Bs:
    env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", role_warps=rw); env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 6 if B <= 65536 else 3
    e0.record()
    for _ in range(n): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{tag} role_warps={rw} B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s", flush=True)
    env.close()
