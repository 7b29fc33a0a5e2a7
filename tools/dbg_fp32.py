import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
from glgym.weather import load_weather_data
tabs = np.stack([load_weather_data(None, "Bleiswijk", "GL", 2009, sd, 60, 49, 900, 10) for sd in (0, 5, 12)])
for name, kw in (("noise0.3", dict(uncertainty_scale=0.3, seed=42)),):
    B = 96
    e64, e32 = GreenLightVecEnv(B, n_sub=600, precision="fp64", **kw), GreenLightVecEnv(B, n_sub=600, precision="fp32", **kw)
    e64.reset_tensor(); e32.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for s in range(20):
        a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
        e64.step_tensor(a); e32.step_tensor(a)
        if s in (0, 1, 19):
            x64, x32 = e64.state_t.cpu().numpy(), e32.state_t.cpu().numpy()
            err = np.abs(x32 - x64) / np.maximum(np.abs(x64), 1e-3)
            print(name, "step", s + 1, "max rel err", err.max(), "state", np.unravel_index(err.argmax(), err.shape), "per-state", np.array2string(err.max(axis=1)[[0,4,15,22,23,25]], precision=1))
    e64.close(); e32.close()
