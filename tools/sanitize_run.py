"""Small runs of every kernel variant for compute-sanitizer (memcheck / racecheck): both layouts of the step kernel, kernel A, both
integrator contracts, fp32 units, parametric uncertainty, the on-device rule-based controller, a non-default observation stack
(no forecast block / StateObservations), the overlapped and the split host path, the device rollout kernels (store, GAE, carry), termination with
in-place reset.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py ; compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
import numpy as np
import torch
from glgym.rollout import DeviceRollout
from glgym.vec_env import GreenLightVecEnv
from glgym.weather import load_weather_data

W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
B = 70
short = dict(season_length=3 / 96.0)  # episodes end after 4 steps: auto-reset + terminal observation paths run
for kw in (dict(integrator="fixed", role_warps=2), dict(integrator="fixed", role_warps=3), dict(integrator="graded", n_sub=6, role_warps=2),
           dict(integrator="graded", n_sub=6, role_warps=3), dict(integrator="fixed", uncertainty_scale=0.3), dict(integrator="fixed", precision="fp32"),
           dict(integrator="fixed", role_warps=1),
           dict(integrator="fixed", observation_modules=["StateObservations", "ControlObservations", "IndoorClimateObservations"]),
           dict(integrator="fixed", observation_modules=["WeatherForecastObservations", "BasicCropObservations"])):
    kw.setdefault("n_sub", 8)
    env = GreenLightVecEnv(B, base_env_params=short, weather_tables=W, **kw)
    env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(5):
        env.step_tensor(A)
    if kw.get("role_warps") != 1:
        env.step_rule_based_tensor()
    a = A.cpu().numpy()
    for _ in range(3):  # the first host step has no prediction to work from; the following ones run the overlapped path proper
        env.step(a)
    env.step_split(a)[0].full([0, B - 1])
    roll = DeviceRollout(env, 3, gamma=0.96, gae_lambda=0.9)
    roll.reset()
    for _ in range(3):
        roll.step(A)
    roll.finish(torch.zeros(4, B, device="cuda"))
    roll.begin()
    sd = env.state_dict()
    env.load_state_dict(sd)
    torch.cuda.synchronize()
    print("ok", kw, bool(torch.isfinite(env.obs_t).all()), flush=True)
    env.close()
