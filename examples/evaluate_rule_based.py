"""The reference's stochastic rule-based evaluation sweep (experiments/eval_baseline.sh: uncertainty scales 0 ... 0.3, 30
simulations each, one season per simulation) on one GPU: the 30 simulations of a scale are the envs of one batch, each with its
own Philox noise stream; the controller runs on the device.  Writes one CSV per scale in the reference's Results layout
(experiments/evaluate_baseline.py:63-67), readable by gl_gym/visualisations.

    python examples/evaluate_rule_based.py [--out data/rb_baseline] [--sims 30] [--integrator fixed|graded]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym.vec_env import GreenLightVecEnv
from glgym.evaluation import evaluate_rule_based, result_columns, save_results

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="data/rb_baseline")
ap.add_argument("--sims", type=int, default=30)
ap.add_argument("--scales", type=float, nargs="*", default=[0.0, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3])
ap.add_argument("--integrator", default="fixed")
ap.add_argument("--steps", type=int, default=None, help="steps per simulation (default: one season, N + 1)")
a = ap.parse_args()
os.makedirs(a.out, exist_ok=True)
t_all = time.time()
for scale in a.scales:
    t0 = time.time()
    env = GreenLightVecEnv(a.sims if scale > 0 else 1, uncertainty_scale=scale, seed=666, integrator=a.integrator, info_mode="minimal")
    data = evaluate_rule_based(env, None, n_steps=a.steps)      # [sims, steps, 32]
    path = os.path.join(a.out, f"rb_baseline_scale{scale:g}.csv")
    save_results(data, result_columns(env), path)
    print(f"scale {scale:g}: {data.shape[0]} simulations x {data.shape[1]} steps in {time.time() - t0:.1f} s, "
          f"mean episode return {data[:, :, 23].sum(axis=1).mean():.3f} -> {path}", flush=True)
    env.close()
print(f"total {time.time() - t_all:.1f} s")
