"""Divergence of the fast modes from the fp64 exact path over 500 free-running steps, turbulence on (the numbers behind
the bounds asserted in tests/test_gpu_env.py::test_fast_modes_at_their_stated_tolerance).  python tools/fast_mode_divergence.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config

n = 2048
for turb in (False, True):
    kw = dict(sim_config_kw={"turbulence": turb}, seed=5)
    for precision, integrator in (("f32", "rk45"), ("f64", "rk4"), ("f32", "rk4")):
        ref = bt.BatchedFixedWing(n, cfg=build_config(**kw))
        fast = bt.BatchedFixedWing(n, cfg=build_config(precision=precision, integrator=integrator, rk4_substeps=4, **kw))
        for e in (ref, fast):
            e.enable_f64_outputs()
            e.reset()
        rs = np.random.RandomState(1)
        alive = np.ones(n, bool)
        for t in range(500):
            ph = 0.02 * t
            a = torch.as_tensor(np.stack([0.3 * np.sin(ph + rs.rand()) * np.ones(n), 0.3 * np.cos(ph) * np.ones(n),
                                          0.5 * np.ones(n)], 1).astype(np.float32)).cuda()
            ref.step(a, auto_reset=False)
            fast.step(a, auto_reset=False)
            alive &= ~(ref.done.cpu().numpy().astype(bool) | fast.done.cpu().numpy().astype(bool))
            if t in (99, 249, 499):
                o0, o1 = ref.obs64.cpu().numpy(), fast.obs64.cpu().numpy()
                d = np.abs(o1 - o0)[alive][:, :6].max(axis=1)
                print("turb=%d %s/%s step %3d alive %4d: median %.2e p90 %.2e p99 %.2e max %.2e" % (
                    turb, precision, integrator, t + 1, alive.sum(), np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()))
        ref.close(); fast.close()
