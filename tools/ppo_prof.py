import sys, time; sys.path.insert(0,'.')
import torch
from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
venv = FixedWingVecEnv(8192, sim_config_kw={"turbulence": True}, seed=0)
import sys as _s
algo = PPO(venv, n_steps=32, batch_size=32768, n_epochs=10, use_cuda_graph=("--eager" not in _s.argv))
algo._setup()
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    algo.collect_rollouts(); torch.cuda.synchronize(); t1=time.perf_counter()
    algo.train(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("rollout %.1f ms  train %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3), flush=True)
# env step alone at 8192
a=torch.rand(8192,3,device='cuda')*2-1
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(100): venv.step_tensor(a)
torch.cuda.synchronize(); print("env step 8192: %.3f ms"%((time.perf_counter()-t0)*10))
with torch.no_grad():
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(100): algo.policy(algo._last_obs)
    torch.cuda.synchronize(); print("policy fwd: %.3f ms"%((time.perf_counter()-t0)*10))
