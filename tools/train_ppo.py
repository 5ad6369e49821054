"""Trains PPO on the default attitude task (config C4: 8192 envs per GPU) for a wall-clock budget and writes the
learning curve as JSON.  Usage: python tools/train_ppo.py [seconds] [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/ppo_curve.json"
hp = dict(n_steps=32, batch_size=65536, n_epochs=10, ent_coef=0.01, learning_rate=3e-4)   # ent_coef, 4 minibatches as SB2 PPO2
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    hp[k] = type(hp[k])(v)
print("hyper-parameters", hp, flush=True)
venv = FixedWingVecEnv(8192, sim_config_kw={"turbulence": True}, seed=0)
algo = PPO(venv, **hp)
t0 = time.time()
rows = []


def cb(row):
    row["wall_s"] = time.time() - t0
    rows.append(row)
    if row["iteration"] % 100 == 0:
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)


chunk = 20 * 32 * 8192
while time.time() - t0 < budget:
    algo.learn(total_timesteps=algo.num_timesteps + chunk, callback=cb)
torch.cuda.synchronize()
train_wall = time.time() - t0

# ---- evaluation: deterministic policy, one full episode (2000 steps) on 8192 fresh envs, the env's own metrics ----
import numpy as np
from tum_adlr_deep_reinforcement_learning_b200.config import METRIC_LAYOUT
ev = FixedWingVecEnv(8192, sim_config_kw={"turbulence": True}, seed=12345)
obs = algo.norm.normalize_obs(ev.reset_tensor())
ret = torch.zeros(8192, dtype=torch.float64, device=ev.device)
rows_m = []
with torch.no_grad():
    for t in range(ev.cfg.steps_max):
        a, _, _ = algo.policy(obs, deterministic=True)
        o, r, d = ev.step_tensor(a.contiguous())
        ret += r.double()
        if bool(d.any()):
            term, m, er, el = ev.sim.episode_info()
            sel = d.bool()
            rows_m.append(torch.cat([m[sel], er[sel, None], el[sel, None].double(), term[sel, None].double()], 1).cpu().numpy())
        obs = algo.norm.normalize_obs(o)
M = np.concatenate(rows_m)
evalres = {"episodes": int(M.shape[0]), "ep_rew_mean": float(M[:, 28].mean()), "ep_len_mean": float(M[:, 29].mean()),
           "failures": int((M[:, 30] >= 10).sum())}
for name, off, keys in METRIC_LAYOUT:
    if name in ("success", "success_time_frac", "settling_time", "rise_time", "control_variation", "overshoot"):
        evalres[name] = {k: float(np.nanmean(M[:, off + j])) for j, k in enumerate(keys)}
print("EVAL", json.dumps(evalres), flush=True)
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"config": {"envs": 8192, **hp, "turbulence": "light",
                      "mode": "fp64 exact"}, "wall_s": train_wall, "timesteps": algo.num_timesteps, "eval": evalres,
           "curve": [r for r in rows if "iteration" in r]}, open(out, "w"))
print("done", algo.num_timesteps, "steps in", round(train_wall, 1), "s")
