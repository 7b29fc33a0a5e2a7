"""Rule-based evaluation (BASELINE config 1's shape): host controller + step_raw_control vs the fused device controller."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
from glgym.controller import RuleBasedController
ctrl = RuleBasedController()
for B, steps in ((1, 400), (30, 400), (4096, 60)):
    env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", info_mode=None); env.reset()
    W = env.weather_tables[0] if hasattr(env, "weather_tables") else None
    t0 = time.perf_counter()
    for s in range(steps):
        x, u, k = env.get_state()
        tm = env.time_t.cpu().numpy() if hasattr(env, "time_t") else None
        d = env.get_attr("weather_data")[0][k] if W is None else W[k]
        uu = ctrl.predict(x, d, tm[1] if tm is not None else 0.0, tm[0] if tm is not None else 0.0)
        env.step_raw_control(uu)
    host = B * steps / (time.perf_counter() - t0)
    env.close()
    env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", info_mode=None); env.reset_tensor()
    for _ in range(3): env.step_rule_based_tensor()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(steps): env.step_rule_based_tensor()
    torch.cuda.synchronize(); dev = B * steps / (time.perf_counter() - t0)
    print(f"B={B}: host controller loop {host:.3e} env-steps/s ; fused device controller {dev:.3e} env-steps/s ({dev / host:.1f}x)", flush=True)
    env.close()
