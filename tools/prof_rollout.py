"""ncu target for the device rollout kernels (csrc/glg_rollout.cuh) at 65 536 envs: one glg_rollout_store (moments, finish, apply) and
one glg_rollout_gae behind a reset + step.   ncu -k regex:glg_roll --launch-skip 6 --launch-count 4 python tools/prof_rollout.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "greenlight-gym2_b200"))
import torch
from glgym.vec_env import GreenLightVecEnv
from glgym.rollout import DeviceRollout
B = 65536
env = GreenLightVecEnv(B); roll = DeviceRollout(env, 4)
roll.reset(); roll.step(torch.zeros(B, 6, device="cuda"))
roll._store(1)
roll.finish(torch.randn(5, B, device="cuda"))
torch.cuda.synchronize()
