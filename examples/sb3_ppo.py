"""Training the reference's PPO agent on the batched GPU env (needs stable-baselines3, which this image does not ship):
`GreenLightVecEnv` speaks SB3's VecEnv protocol, so the reference's wrappers and hyper-parameters apply unchanged
(gl_gym/RL/utils.py:44-69, gl_gym/configs/agents/ppo.yml).

    python examples/sb3_ppo.py --n-envs 4096 --total-timesteps 20000000
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--n-envs", type=int, default=4096)
ap.add_argument("--total-timesteps", type=int, default=20_000_000)
ap.add_argument("--uncertainty-scale", type=float, default=0.0)
a = ap.parse_args()
try:
    import torch
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env import VecMonitor, VecNormalize
except ImportError as exc:
    sys.exit(f"stable-baselines3 is required for this example ({exc})")
from glgym.vec_env import GreenLightVecEnv

env = GreenLightVecEnv(a.n_envs, uncertainty_scale=a.uncertainty_scale, seed=666)
env = VecNormalize(VecMonitor(env), norm_obs=True, norm_reward=True, clip_obs=10.0, gamma=0.9631)
model = PPO("MlpPolicy", env, n_steps=64, batch_size=4096, n_epochs=8, gamma=0.9631, gae_lambda=0.9167, clip_range=0.2,
            ent_coef=0.05434, vf_coef=0.8225, max_grad_norm=0.3, learning_rate=2e-5,
            policy_kwargs=dict(net_arch=dict(pi=[256] * 3, vf=[512] * 3), activation_fn=torch.nn.SiLU, log_std_init=0.0),
            device="cuda")
model.learn(total_timesteps=a.total_timesteps)
model.save("ppo_greenlight")
