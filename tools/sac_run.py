"""SAC on config C5 (1024 envs, turbulence, replay ring in HBM) for a wall-clock budget: env-steps/s and reward."""
import sys, time, json; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200.sac import SAC
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
env = FixedWingVecEnv(1024, sim_config_kw={"turbulence": True}, seed=0)
algo = SAC(env, buffer_size=1_000_000, batch_size=4096, gradient_steps=2, learning_starts=10_000)
t0 = time.time(); last = {}
def cb(row):
    last.update(row)
while time.time() - t0 < budget:
    algo.learn(total_timesteps=algo.num_timesteps + 200 * 1024, log_every=50, callback=cb)
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps({"env_steps": algo.num_timesteps, "wall_s": dt, "env_steps_per_s": algo.num_timesteps / dt,
                  "gradient_steps_per_s": 2 * (algo.num_timesteps - 10_000) / 1024 / dt, "last": last}))
