"""Quick GPU check: device math accuracy, step parity vs oracle (few envs), timings. Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from glgym import _lib
from glgym.params import init_default_params
from glgym.weather import load_weather_data
from glgym.vec_env import GreenLightVecEnv
import oracle_binding as ob
L = _lib.load()
rng = np.random.default_rng(0)
def dev_math(op, x):
    xi = torch.as_tensor(x, device="cuda"); yo = torch.empty_like(xi)
    assert L.glg_debug_math(op, xi.data_ptr(), yo.data_ptr(), xi.numel(), 0) == 0
    torch.cuda.synchronize(); return yo.cpu().numpy()
x = rng.uniform(-700, 700, 1 << 20); print("exp rel", np.max(np.abs(dev_math(0, x) / np.exp(x) - 1)))
x = np.exp(rng.uniform(-40, 40, 1 << 20)); print("log abs", np.max(np.abs(dev_math(1, x) - np.log(x))))
print("rcp rel", np.max(np.abs(dev_math(2, x) * x - 1)), "sqrt rel", np.max(np.abs(dev_math(3, x) / np.sqrt(x) - 1)))
x = np.exp(rng.uniform(-25, 6, 1 << 20)); print("cbrt rel", np.max(np.abs(dev_math(4, x) / np.cbrt(x) - 1)),
      "pow.66", np.max(np.abs(dev_math(5, x) / x ** 0.66 - 1)), "pow.32", np.max(np.abs(dev_math(6, x) / x ** 0.32 - 1)))
x = rng.uniform(-800, 800, 1 << 16); print("inv1pexp abs", np.max(np.abs(dev_math(7, x) - 1 / (1 + np.exp(np.clip(x, -708, 709))))))
print("edge exp", dev_math(0, np.array([-1e4, 1e4, 0.0, np.nan])), "cbrt0", dev_math(4, np.array([0.0, 1e-300])))

p = init_default_params().astype(np.float64)
W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
kw = {}
if len(sys.argv) > 1: kw["role_warps"] = int(sys.argv[1])
B = 96
env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", **kw); env.reset()
orc = [ob.OracleEnv(W, p) for _ in range(0, B, 5)]
for o in orc: o.reset()
for s in range(3):
    A = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
    obs, rew, done, infos = env.step(A); xg, ug, kg = env.get_state()
    ws = wo = wr = 0
    for j, b in enumerate(range(0, B, 5)):
        o, r, dn, info = orc[j].step(action=A[b])
        ws = max(ws, (np.abs(xg[b] - orc[j].x) / np.maximum(np.abs(orc[j].x), 1e-3)).max())
        wo = max(wo, (np.abs(obs[b] - o.astype(np.float32)) / np.maximum(np.abs(o), 1e-3)).max()); wr = max(wr, abs(rew[b] - r))
    print(f"step {s}: state rel {ws:.2e} obs rel {wo:.2e} reward abs {wr:.2e}")
env.close()
for B in (4096, 16384, 65536, 262144):
    env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", **kw); env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 4; e0.record()
    for _ in range(n): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / n
    print(f"B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s  frac_of_34.2TF={B / ms * 1e3 * 2.381e6 / 34.2e12:.3f}")
    env.close()
