"""First GPU contact: parity of evalF + fused step against the oracle and raw timings. Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np, torch
from glgym import _lib
from glgym.params import init_default_params
from glgym.weather import load_weather_data, init_state
from glgym.model import GreenLight
from glgym.vec_env import GreenLightVecEnv
import oracle_binding as ob

torch.cuda.init()
print(torch.cuda.get_device_name(0))
L = _lib.load()
pk = C.c_double()
print("fp64 peak rc", L.glg_measure_fp64_peak(0, C.byref(pk)), pk.value / 1e12, "TFLOP/s")
print("fp32 peak rc", L.glg_measure_fp32_peak(0, C.byref(pk)), pk.value / 1e12, "TFLOP/s")

p = init_default_params().astype(np.float64)
W = load_weather_data(None, "Bleiswijk", "GL", 2009, 0, 60, 49, 900, 10)
rng = np.random.default_rng(0)
# ---- evalF parity
B = 256
x0 = init_state(W[0])
X = np.tile(x0, (B, 1)); X[:, 2:15] += rng.uniform(-5, 8, (B, 13)); X[:, 15:17] *= rng.uniform(0.6, 1.2, (B, 2))
U = rng.uniform(0, 1, (B, 6)); D = W[rng.integers(0, 5000, B)]
gl = GreenLight(n_sub=600)
t = time.time(); Y = gl.evalF_batch(X, U, D, p).cpu().numpy(); print("gpu evalf", time.time() - t)
t = time.time(); Yo = ob.evalf_batch(X, U, D, p, n_sub=600); print("cpu evalf", time.time() - t)
rel = np.abs(Y - Yo) / np.maximum(np.abs(Yo), 1e-3)
print("evalF max rel err", rel.max(), "state", np.unravel_index(rel.argmax(), rel.shape))
# per-env p
PP = np.tile(p, (32, 1)); PP[:, 128:162] *= 1 + rng.uniform(-.1, .1, (32, 34))
Y2 = gl.evalF_batch(X[:32], U[:32], D[:32], PP).cpu().numpy(); Yo2 = ob.evalf_batch(X[:32], U[:32], D[:32], PP, n_sub=600)
print("evalF per-env-p max rel", (np.abs(Y2 - Yo2) / np.maximum(np.abs(Yo2), 1e-3)).max())

# ---- fused step parity (B=96 -> one full + one partial block), 6 steps
B = 96
env = GreenLightVecEnv(B, n_sub=600)
obs0 = env.reset()
orc = [ob.OracleEnv(W, p) for _ in range(B)]
o0 = orc[0].reset()
print("reset obs max abs diff", np.abs(obs0[0] - o0.astype(np.float32)).max())
for s in range(4):
    A = rng.uniform(-1, 1, (B, 6)).astype(np.float32)
    obs, rew, done, infos = env.step(A)
    xg, ug, kg = env.get_state()
    worst = 0; worst_o = 0; worst_r = 0
    for b in range(0, B, 7):
        o, r, dn, info = orc[b].step(action=A[b])
        worst = max(worst, (np.abs(xg[b] - orc[b].x) / np.maximum(np.abs(orc[b].x), 1e-3)).max())
        worst_o = max(worst_o, (np.abs(obs[b] - o.astype(np.float32)) / np.maximum(np.abs(o), 1e-3)).max())
        worst_r = max(worst_r, abs(rew[b] - r))
    print(f"step {s}: state rel {worst:.2e} obs rel {worst_o:.2e} reward abs {worst_r:.2e} done {done.sum()}")

# ---- timing
for B in (4096, 65536):
    env = GreenLightVecEnv(B, n_sub=600)
    env.reset_tensor()
    A = torch.rand(B, 6, device="cuda") * 2 - 1
    for _ in range(2): env.step_tensor(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): env.step_tensor(A)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} env-steps/s  frac_of_37.2TF={B / ms * 1e3 * 2.381e6 / 37.2e12:.3f}")
    env.close()
