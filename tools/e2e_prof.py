import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
n = 65536
v = FixedWingVecEnv(n, sim_config_kw={"turbulence": True}, seed=0)
v.reset()
a = np.random.uniform(-1, 1, (n, 3)).astype(np.float32)
for _ in range(10): v.step(a)
def t(f, k=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
print("full step            %.0f us" % t(lambda: v.step(a)))
print("act host copy+nan    %.0f us" % t(lambda: (v._act_np.__setitem__(Ellipsis, a), np.isnan(v._act_np.sum()))))
print("H2D actions          %.0f us" % t(lambda: (v._act_dev.copy_(v._act_pin, non_blocking=True), torch.cuda.current_stream().synchronize())))
print("sim.step (3 kernels) %.0f us" % t(lambda: (v.sim.step(v._act_dev), torch.cuda.current_stream().synchronize())))
o, r, d = v._out[0]
print("D2H obs              %.0f us" % t(lambda: (o.copy_(v.sim.obs, non_blocking=True), torch.cuda.current_stream().synchronize())))
print("D2H rew+done         %.0f us" % t(lambda: (r.copy_(v.sim.rew, non_blocking=True), d.copy_(v.sim.done, non_blocking=True), torch.cuda.current_stream().synchronize())))
dn = v._out_np[0][2]
print("build infos (no done)%.0f us" % t(lambda: v._build_infos(dn)))
