"""Does writing obs | rew | done straight into pinned host memory from head_kernel beat kernel + cudaMemcpyAsync D2H?"""
import sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, ".")
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt, _lib
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
n = 65536
env = bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True}))
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
acts = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1) for _ in range(8)]
host = torch.zeros(env.out_nbytes, dtype=torch.uint8).pin_memory()
hobs, hrew, hdone = bt.unpack_outputs(host, n, env.obs_dim)
pin2 = torch.zeros(env.out_nbytes, dtype=torch.uint8).pin_memory()
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def step_copy(a):
    env.step(a)
    pin2.copy_(env.out_packed, non_blocking=True)
    torch.cuda.current_stream().synchronize()

def step_zero(a):
    _lib.check(_lib.lib().fw_step(env._h, p(a), 0, p(hobs), p(hrew), p(hdone), p(env.term_obs), None, None, 1, st()), "fw_step")
    torch.cuda.current_stream().synchronize()

for name, fn in (("kernel + D2H copy", step_copy), ("zero-copy stores", step_zero), ("kernel + D2H copy", step_copy), ("zero-copy stores", step_zero)):
    for k in range(10):
        fn(acts[k % 8])
    t0 = time.perf_counter()
    for k in range(100):
        fn(acts[k % 8])
    dt = (time.perf_counter() - t0) / 100
    print("%-20s %.1f us per step" % (name, dt * 1e6))
o1 = hobs.clone(); step_copy(acts[0])
print("host obs finite:", bool(torch.isfinite(hobs).all()), float(hobs.abs().sum()))
