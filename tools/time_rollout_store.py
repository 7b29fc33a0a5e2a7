import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
from glgym.rollout import DeviceRollout
for B in (4096, 65536):
    env = GreenLightVecEnv(B); roll = DeviceRollout(env, 4)
    roll.reset(); roll.step(torch.zeros(B, 6, device="cuda"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): roll._store(1)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: glg_rollout_store {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call ({3 * B * 263 * 4 / 1e6:.0f} MB of algorithmic traffic)")
    v = torch.randn(5, B, device="cuda")
    e0.record()
    for _ in range(20): roll.finish(v)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: glg_rollout_gae (T=4) {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
    env.close()
