"""CPU only, build container only (needs /root/reference): a randomised sweep of the oracle against the LIVE reference
beyond the committed fixtures.  Fresh seeds, every turbulence intensity, calm / windy / hard initial states, large
action amplitudes — and then every config file the reference ships (general observation layouts, reward variants);
every step compares what tests/test_oracle_golden.py compares (done / termination / RK45 RHS count
exact, state / observation / reward <= 1e-9 relative).  Prints the worst deviation per category and every mismatch.

    python tools/oracle_sweep.py [episodes_per_category] [steps] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as G  # noqa: E402  (installs oracle/refshim and imports the reference)
from oracle import fw_oracle as O  # noqa: E402
from tum_adlr_deep_reinforcement_learning_b200.config import build_config  # noqa: E402

N_EP = int(sys.argv[1]) if len(sys.argv) > 1 else 12
N_STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 20260101

CASES = [  # name, turbulence, intensity, wind_mag, action_amp, hard, f32 actions
    ("calm", False, "light", 0.0, 1.0, False, False),
    ("calm_clip", False, "light", 0.0, 2.5, False, False),
    ("wind", False, "light", 8.0, 1.3, False, False),
    ("hard", False, "light", 4.0, 1.5, True, False),
    ("f32act", False, "light", 3.0, 1.2, False, True),
    ("turb_light", True, "light", 8.0, 1.0, False, False),
    ("turb_moderate", True, "moderate", 6.0, 1.3, False, False),
    ("turb_severe", True, "severe", 5.0, 1.5, False, False),
    ("turb_hard", True, "severe", 8.0, 2.0, True, False),
]


# A config without rate constraints (fixed_wing_config_dev.json, scale_space false) lets the model blow up in finite time:
# in the reference the body rates pass 100 rad/s, then the state climbs from 1e4 to 1e154 within a dozen steps (16 000 RHS
# evaluations per step — the oracle counts the same 16 580 —, numpy overflow warnings).  Round-off differences are
# amplified without bound there, so an episode is compared up to the step at which the REFERENCE state (position aside)
# exceeds BLOWUP and counted in `blown`; up to that step the usual bounds hold.
BLOWUP = 200.0
blown = 0


def rel(a, b):
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


bad = 0
for ci, (name, turb, inten, wind, amp, hard, f32) in enumerate(CASES):
    t0 = time.time()
    rs = np.random.RandomState(SEED + ci)
    env = G.make_env(turb, intensity=inten)
    g = G.run_episodes(env, N_EP, N_STEPS, rs, turb, wind_mag=wind, action_amp=amp, hard=hard, f32_actions=f32)
    cfg = build_config(sim_config_kw={"turbulence": turb, "turbulence_intensity": inten})
    worst = {"y": 0.0, "obs": 0.0, "rew": 0.0}
    steps = ends = 0
    for ep in range(N_EP):
        o = O.OracleEnv(cfg)
        obs = o.reset(g["init_state"][ep], g["init_target"][ep], g["noise"][ep] if turb else None)
        if np.abs(obs - g["obs0"][ep]).max() > 1e-12:
            print("  MISMATCH", name, ep, "reset observation"); bad += 1
        for t in range(int(g["n_valid"][ep])):
            if np.abs(np.delete(g["y"][ep, t], [7, 8, 9])).max() > BLOWUP:      # position excluded
                blown += 1
                break
            obs, rew, done, term = o.step(g["actions"][ep, t], f32)
            s = o.get()
            steps += 1
            if done != bool(g["done"][ep, t]) or term != int(g["term"][ep, t]):
                print("  MISMATCH", name, ep, t, "done/term", done, term, g["done"][ep, t], g["term"][ep, t]); bad += 1
                break
            if s["nfev"] != int(g["nfev"][ep, t]):
                print("  MISMATCH", name, ep, t, "nfev", s["nfev"], g["nfev"][ep, t]); bad += 1
            if term < 10:
                worst["y"] = max(worst["y"], rel(s["y"], g["y"][ep, t]))
            worst["obs"] = max(worst["obs"], rel(obs, g["obs"][ep, t]))
            worst["rew"] = max(worst["rew"], abs(rew - g["reward"][ep, t]) / max(1.0, abs(g["reward"][ep, t])))
            ends += int(done)
    flag = "" if max(worst.values()) < 1e-9 else "  <-- above 1e-9"
    bad += int(bool(flag))
    print("%-14s %5d steps, %2d episode ends, worst rel: y %.1e obs %.1e reward %.1e  (%.0f s)%s"
          % (name, steps, ends, worst["y"], worst["obs"], worst["rew"], time.time() - t0, flag), flush=True)

# ---- every config file the reference ships (general observation layouts, reward variants), by path ----
# The env-level random draws of the reference (the init_noise offset of the observation history, fixed_wing.py:1145,
# and the observation noise) are pinned with make_golden.FixedDraws (u = 0.25, noise = its mean) — the oracle gets the
# same offset and a zero noise variance; matrix observations are compared flattened (row-major).
import glob  # noqa: E402
import json  # noqa: E402

base = os.path.normpath(os.path.join(os.path.dirname(G.refshim.GYM_CONFIG), "..", "..", ".."))
files = sorted(glob.glob(os.path.join(base, "**", "fixed_wing_config*.json"), recursive=True))
n_cfg = 0
for fi, path in enumerate(files):
    try:
        raw = json.load(open(path))
    except json.JSONDecodeError:
        continue                                   # fixed_wing_config-commented.json
    okw = {"observation": {"noise": {"mean": 0, "var": 0}}} if "noise" in raw["observation"] else None
    turb = fi % 2 == 1
    sim_kw = {"turbulence": turb, "turbulence_intensity": "moderate"}
    try:
        cfg = build_config(env_cfg=path, config_kw=okw, sim_config_kw=sim_kw, obs_init_noise=0.25)
    except NotImplementedError:
        continue                                   # archived position-target experiments: refused loudly
    t0 = time.time()
    inner = G.make_env(turb, intensity="moderate", config_path=path)
    inner.np_random = G.FixedDraws()

    class Flat:
        def __getattr__(self, k):
            return getattr(inner, k)

        def reset(self, **kw):
            return np.asarray(inner.reset(**kw)).ravel()

        def step(self, a):
            o_, r_, d_, i_ = inner.step(a)
            return np.asarray(o_).ravel(), r_, d_, i_

    rs = np.random.RandomState(SEED + 100 + fi)
    n_ep = max(2, N_EP // 3)
    g = G.run_episodes(Flat(), n_ep, N_STEPS, rs, turb, wind_mag=5.0, action_amp=1.3)
    worst = {"y": 0.0, "obs": 0.0, "rew": 0.0}
    steps = 0
    for ep in range(n_ep):
        o = O.OracleEnv(cfg)
        obs = o.reset(g["init_state"][ep], g["init_target"][ep], g["noise"][ep] if turb else None)
        if np.abs(obs - g["obs0"][ep]).max() > 1e-12:
            print("  MISMATCH", path, ep, "reset observation"); bad += 1
        for t in range(int(g["n_valid"][ep])):
            if np.abs(np.delete(g["y"][ep, t], [7, 8, 9])).max() > BLOWUP:      # position excluded
                blown += 1
                break
            obs, rew, done, term = o.step(g["actions"][ep, t], False)
            s = o.get()
            steps += 1
            if done != bool(g["done"][ep, t]) or term != int(g["term"][ep, t]) or s["nfev"] != int(g["nfev"][ep, t]):
                print("  MISMATCH", path, ep, t, "done / term / nfev"); bad += 1
                break
            if term < 10:
                worst["y"] = max(worst["y"], rel(s["y"], g["y"][ep, t]))
            worst["obs"] = max(worst["obs"], rel(obs, g["obs"][ep, t]))
            worst["rew"] = max(worst["rew"], abs(rew - g["reward"][ep, t]) / max(1.0, abs(g["reward"][ep, t])))
    flag = "" if max(worst.values()) < 1e-9 else "  <-- above 1e-9"
    bad += int(bool(flag))
    n_cfg += 1
    print("%-58s turb %d  %5d steps, worst rel: y %.1e obs %.1e reward %.1e  (%.0f s)%s"
          % (os.path.relpath(path, base)[-58:], turb, steps, worst["y"], worst["obs"], worst["rew"], time.time() - t0, flag),
          flush=True)
print("config files replayed:", n_cfg)
print("episodes cut at a numerical blow-up of the reference (|y| > %g):" % BLOWUP, blown)
print("MISMATCHES:", bad)
sys.exit(1 if bad else 0)
