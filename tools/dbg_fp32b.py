import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
from glgym.params import PARAM_NAMES
B = 68
noise = np.zeros((B, 34)); 
for i in range(34): noise[i, i] = 0.15; noise[34 + i, i] = -0.15
n = torch.as_tensor(noise, device="cuda")
e64, e32 = GreenLightVecEnv(B, n_sub=600, precision="fp64", uncertainty_scale=0.3), GreenLightVecEnv(B, n_sub=600, precision="fp32", uncertainty_scale=0.3)
e64.reset_tensor(); e32.reset_tensor()
a = torch.zeros(B, 6, device="cuda")
for s in range(3):
    e64.step_tensor(a, noise=n); e32.step_tensor(a, noise=n)
x64, x32 = e64.state_t.cpu().numpy(), e32.state_t.cpu().numpy()
err = np.abs(x32 - x64) / np.maximum(np.abs(x64), 1e-3)
for b in np.argsort(-err.max(axis=0))[:8]:
    i = b % 34
    print(f"env {b}: param p[{128+i}] {PARAM_NAMES[128+i]} {'+' if b < 34 else '-'}15%: max rel err {err[:, b].max():.2e} state {err[:, b].argmax()}  x22 {x64[22,b]:.3f} vs {x32[22,b]:.3f}")
