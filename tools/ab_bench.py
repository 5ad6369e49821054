"""A/B of kernel variants: python tools/ab_bench.py build/variants/*.so  (each library in its own process)."""
import json, os, subprocess, sys

CHILD = r'''
import sys, json; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
n = 65536
cfg = build_config(sim_config_kw={"turbulence": True})
env = bt.BatchedFixedWing(n, cfg=cfg); env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
pool = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous() for _ in range(8)]
env.set_profiling(True)
for i in range(30): env.step(pool[i % 8])
fresh = env.profile()
env.set_profiling(False)
for i in range(170): env.step(pool[i % 8])
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
K = 300
e0.record()
for i in range(K): env.step(pool[i % 8])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
env.set_profiling(True)
for i in range(100): env.step(pool[i % 8])
p = env.profile()
cs = float(env.obs.double().sum()) + float(env.get_field(bt.FIELD_Y).sum())
print(json.dumps(dict(ms=ms, rate=n / ms * 1e3, init_us=p["init_ms"] * 1e3, integ_us=p["integrate_ms"] * 1e3,
                      head_us=p["head_ms"] * 1e3, fresh_init_us=fresh["init_ms"] * 1e3,
                      fresh_head_us=fresh["head_ms"] * 1e3, done_rate=float(env.done.float().mean()), checksum=cs)))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, FWB200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip() or r.stderr[-400:], flush=True)
