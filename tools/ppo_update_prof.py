"""Kernel list of one PPO minibatch update (eager) via torch.profiler."""
import sys; sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
venv = FixedWingVecEnv(8192, sim_config_kw={"turbulence": True}, seed=0)
algo = PPO(venv, n_steps=32, batch_size=32768, n_epochs=1, use_cuda_graph=False)
algo._setup()
algo.collect_rollouts(); algo.train()
from tum_adlr_deep_reinforcement_learning_b200.buffers import RolloutBufferSamples
flat = [algo.buffer.flat(x) for x in (algo.buffer.observations, algo.buffer.actions, algo.buffer.values,
                                      algo.buffer.log_probs, algo.buffer.advantages, algo.buffer.returns)]
batch = RolloutBufferSamples(*(f[:32768].contiguous() for f in flat))
for _ in range(3): algo._minibatch_update(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo._minibatch_update(batch)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
tot = sum(e.device_time for e in ev)
print("kernels", len(ev), "total device us", tot)
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:70]][0] += 1; agg[e.name[:70]][1] += e.device_time
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%4d %8.1f us  %s" % (c, t, k))
