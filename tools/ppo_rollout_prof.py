"""Kernel list of one PPO rollout step (eager) via torch.profiler."""
import sys; sys.path.insert(0, ".")
import collections, torch
from torch.profiler import profile, ProfilerActivity
from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
venv = FixedWingVecEnv(8192, sim_config_kw={"turbulence": True}, seed=0)
algo = PPO(venv, n_steps=32, batch_size=32768, n_epochs=1, use_cuda_graph=False)
algo._setup()
algo.collect_rollouts()
algo.buffer.reset()
for t in range(3): algo._rollout_step(t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo._rollout_step(3)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
print("kernels", len(ev), "total device us", sum(e.device_time for e in ev))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:80]][0] += 1; agg[e.name[:80]][1] += e.device_time
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%4d %8.1f us  %s" % (c, t, k))
