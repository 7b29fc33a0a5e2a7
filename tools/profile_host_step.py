"""Where the host-side time of GreenLightVecEnv.step(numpy) goes: cProfile over 400 steps at B = 4096 (run under gpurun)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "greenlight-gym2_b200"))
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
B = 4096
env = GreenLightVecEnv(B); env.reset()
a = np.random.default_rng(0).uniform(-1, 1, (8, B, 6)).astype(np.float32)
for i in range(5): env.step(a[i % 8])
for name, fn in (("step", env.step), ("step_split", env.step_split)):
    t0 = time.perf_counter()
    for i in range(200): fn(a[i % 8])
    dt = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    at = torch.as_tensor(a[0], device="cuda")
    e0.record()
    for i in range(50): env.step_tensor(at)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {dt * 1e6:.0f} us per call; kernel alone {e0.elapsed_time(e1) / 50 * 1e3:.0f} us")
pr = cProfile.Profile(); pr.enable()
for i in range(400): env.step(a[i % 8])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
