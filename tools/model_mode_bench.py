"""Step rate with per-episode aircraft-parameter randomisation (simulator.model) against the default handle."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import model_env_config
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
n = 65536
g = torch.Generator(device="cuda").manual_seed(1)
pool = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous() for _ in range(8)]
for tag, cfg in (("default", build_config(sim_config_kw={"turbulence": True})),
                 ("simulator.model", build_config(env_cfg=model_env_config(), sim_config_kw={"turbulence": True}))):
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.reset()
    for i in range(20):
        env.step(pool[i % 8])
    env.set_profiling(True)
    for i in range(100):
        env.step(pool[i % 8])
    p = env.profile()
    tot = p["init_ms"] + p["integrate_ms"] + p["head_ms"]
    print("%-16s init %.1f attempt %.1f head %.1f us -> %.3g env-steps/s" % (tag, p["init_ms"] * 1e3, p["integrate_ms"] * 1e3, p["head_ms"] * 1e3, n / tot * 1e3))
    env.close()
