#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched fixed-wing env step (x8 UAV, Dryden turbulence) on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the C port of the reference on every host thread;
                                                            the line also carries the UNMODIFIED Python reference stepped in
                                                            a SubprocVecEnv on the same cores, from baseline/_ref)

Workload (BASELINE.json configs[2], "C3"): 65536 envs PER GPU (weak scaling; envs are independent, no collective
on the step), default attitude task, light Dryden turbulence + steady wind U(-8, 8), actions U(-1,1)^3 that change
every step, fp64 exact mode (bit-for-logic scipy-RK45 replica) — the parity-graded path.  One "step" = one
VecEnv.step over every env of the rank = three launches (rk45_init_kernel, rk45_attempt_kernel, head_kernel).

Timed region (`value`): K launches with the action batches already resident in HBM, one CUDA-event pair per
launch on the launching stream, an L2 flush (256 MiB write) between launches outside the event pairs; barrier +
synchronize on both sides; max over ranks.  `e2e`: the same K steps through FixedWingVecEnv.step(numpy actions)
with the pinned H2D copy of the actions and the D2H copy of obs/reward/done inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (x8 UAV, Dryden turb)"
UNIT = "env-steps/s"
ENVS_PER_GPU = 65536
# algorithmic work per env-step (DESIGN.md "Roofline"; SURVEY §8d hand counts)
W_RHS, W_ATT, W_ENV = 520.0, 1400.0, 400.0           # flops per RHS evaluation / RK45 attempt / env head
T_RHS, T_ATT, T_ENV = 9.0, 1.0, 6.0                  # transcendental calls (counted as 1 flop each as well)
BYTES_PER_ENV_STEP_F64 = 90 * 8 * 2 + 43 * 4 * 2 + 3 * 8 + 12 + 56 + 4 + 1   # SoA state r+w, err ring slot, I/O


def workload_config(n):
    return {"workload": "C3: %d envs/GPU, default attitude task (fixed_wing_config.json), light Dryden turbulence + "
                        "steady wind U(-8,8), fresh U(-1,1)^3 actions every step, auto-reset" % n,
            "envs_per_gpu": n, "mode": "fp64 exact (scipy-RK45 replica, rtol 1e-3 atol 1e-6)",
            "cache": "L2 flushed (256 MiB write) between timed launches; per-launch CUDA events",
            "scaling_unit": "envs sharded by global env id, no data-path collective"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_port_rate(cfg, budget_s, threads):
    """Oracle port (oracle/fw_oracle.c) on the host cores: `threads` python threads each drive a sub-batch through
    ctypes (the GIL is released inside the C call).  Bounded sample; returns (env-steps/s, description)."""
    from oracle import fw_oracle as O
    per = 128
    batches = [O.OracleBatch(cfg_with_offset(cfg, i * per), per) for i in range(threads)]
    for b in batches:
        b.reset()

    def run(b, k, step0):
        b.step_random(k, 1, step0)

    def timed(k, step0):
        ts = [threading.Thread(target=run, args=(b, k, step0)) for b in batches]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    timed(2, 0)                                   # warm-up
    dt = timed(5, 2)
    k = max(5, min(20000, int(budget_s / max(dt / 5, 1e-6))))
    dt = timed(k, 7)
    rate = threads * per * k / dt
    return rate, "%d threads x %d envs x %d steps of the C3 workload (%.1f s)" % (threads, per, k, dt)


def cfg_with_offset(cfg, off):
    import copy
    c = copy.copy(cfg)
    c.env_id_offset = cfg.env_id_offset + off
    return c


def python_reference_worker(budget_s):
    """Runs in a fresh CPU-only process (fork-safe: no CUDA context): the unmodified reference — pyfly + fixed-wing-gym
    under the stable-baselines3 fork's SubprocVecEnv, one worker per host core (BASELINE.md §3 recipe) — imported
    from /root/reference or its mirror baseline/_ref (oracle/refshim.py fabricates the absent gym / matplotlib)."""
    from oracle import refshim
    if not refshim.available():
        print(json.dumps({"unavailable": "reference libraries neither mounted nor mirrored under baseline/_ref"}))
        return
    import warnings
    warnings.filterwarnings("ignore")
    refshim.install()
    from gym_fixed_wing.fixed_wing import FixedWingAircraft
    from stable_baselines3.common.vec_env import SubprocVecEnv
    C = os.cpu_count() or 1

    def make_env(i):
        def f():
            env = FixedWingAircraft(refshim.GYM_CONFIG)
            env.seed(i)
            return env
        return f

    # C1: one env, random actions, reset on done (bounded sample of the 1000-step workload)
    env = make_env(0)()
    env.reset()
    rs = np.random.RandomState(1)
    n1 = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.15 * budget_s:
        _, _, done, _ = env.step(rs.uniform(-1, 1, 3))
        n1 += 1
        if done:
            env.reset()
    single = n1 / (time.perf_counter() - t0)
    # SubprocVecEnv over every core
    venv = SubprocVecEnv([make_env(i) for i in range(C)], start_method="fork")
    venv.reset()
    rs = np.random.RandomState(0)
    for _ in range(5):
        venv.step(rs.uniform(-1, 1, (C, 3)))
    k = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.45 * budget_s or k < 20:
        venv.step(rs.uniform(-1, 1, (C, 3)))
        k += 1
    dt = time.perf_counter() - t0
    # the fork's PPO on the same SubprocVecEnv (BASELINE.md §3 "PPO CPU baseline"), bounded: 2 rollouts of 128 steps per
    # worker instead of SB3's 2048 (the rate is env-bound: 97 % of the wall-clock is env.step), default batch 64, 10 epochs
    ppo_rate, ppo_note = None, None
    try:
        import torch
        from stable_baselines3 import PPO
        torch.set_num_threads(max(1, C // 2))
        n_steps = 128
        model = PPO("MlpPolicy", venv, n_steps=n_steps, batch_size=64, n_epochs=10, device="cpu", verbose=0)
        t1 = time.perf_counter()
        model.learn(total_timesteps=2 * n_steps * C)
        ppo_rate = model.num_timesteps / (time.perf_counter() - t1)
        ppo_note = "PPO('MlpPolicy', SubprocVecEnv x %d, n_steps=%d, batch_size=64, n_epochs=10, device=cpu).learn(%d steps)" % (C, n_steps, model.num_timesteps)
    except Exception as e:                                   # reported, never fatal for the env figure
        ppo_note = "fork PPO baseline failed: %r" % (e,)
    venv.close()
    print(json.dumps({"value": C * k / dt, "unit": UNIT, "cores": C, "workers": C, "kind": "reference",
                      "single_env_value": single, "ppo_train_value": ppo_rate, "ppo_train_sample": ppo_note,
                      "sample": "SubprocVecEnv(start_method=fork) of %d x FixedWingAircraft(fixed_wing_config.json), turbulence "
                                "light, U(-1,1)^3 actions, %d timed vec steps (%.1f s); single env: %d steps" % (C, k, dt, n1)}))


def python_reference_rate(budget_s):
    """The Python reference on the box's host cores, measured in a subprocess; {"unavailable": ...} if it cannot run."""
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "python-reference", "--cpu-budget-s",
                              str(budget_s)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env,
                             timeout=60 + 4 * budget_s)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"unavailable": "python reference run failed: " + (res.stderr.strip().splitlines() or ["?"])[-1][:200]}
        return json.loads(lines[-1])
    except Exception as e:                                   # a baseline that cannot be had is reported, not fatal
        return {"unavailable": repr(e)[:200]}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.  `value` is the C port
    of the reference (oracle/fw_oracle.c) on every host thread — the FASTEST CPU figure we can produce, so the driver's
    ratio is conservative; one timed "step" = every thread advances its 128 envs by 64 env-steps in one C call.  The
    unmodified Python reference (SubprocVecEnv over the same cores) is timed once and reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    cfg = build_config(sim_config_kw={"turbulence": True}, seed=0)
    threads = os.cpu_count() or 1
    from oracle import fw_oracle as O
    per, inner = 128, 64
    batches = [O.OracleBatch(cfg_with_offset(cfg, i * per), per) for i in range(threads)]
    for b in batches:
        b.reset()

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=threads)

    def one_step(step0):
        list(pool.map(lambda b: b.step_random(inner, 1, step0 * inner), batches))

    for w in range(args.warmup):
        one_step(w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(args.warmup + k)
    dt = time.perf_counter() - t0
    n = threads * per * inner
    value = n * args.steps / dt
    sample = "%d threads x %d envs x %d env-steps of the C3 workload per timed step" % (threads, per, inner)
    pyref = python_reference_rate(args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(ENVS_PER_GPU),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "C restatement of the pure-Python reference (the fastest CPU arm available); the "
                                     "reference itself is in python_reference",
                             "python_reference": pyref},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    # NCCL prints its version banner (and any NCCL_DEBUG output) to stdout: keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "python-reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the fp32 / PPO side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference" and args.steps > 400:             # keep the reference arm within a few minutes
        args.steps = 400
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "python-reference":
        return python_reference_worker(args.cpu_budget_s)

    import torch
    import torch.distributed as dist
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the env step has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # the ranks of one box share its host cores: give every rank its own slice, so that the e2e leg (python + pinned
        # copies per step in every process) is not scheduled on top of its neighbours
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local * per:(local + 1) * per] or cores
            os.sched_setaffinity(0, mine)
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, K, W = args.envs_per_gpu, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def time_device_steps(env, pool, flush, n_envs=None):
        """K launches, per-launch events, L2 flush between; returns (total ms of the launches, nfev stats)."""
        n_envs = n if n_envs is None else n_envs
        for w in range(W):
            env.step(pool[w % len(pool)])
        nf_sum = np.zeros(2)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        nf_acc = torch.zeros(2, dtype=torch.float64, device=dev)
        barrier()
        for k in range(K):
            flush.add_(1.0)                      # 256 MiB read+write: evicts L2
            ev[k][0].record()
            env.step(pool[(W + k) % len(pool)])
            ev[k][1].record()
            nf_acc += env.get_field(bt.FIELD_NFEV).to(torch.float64).sum(0)   # outside the event pair
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        nf_sum = (nf_acc / (n_envs * K)).cpu().numpy()
        return ms, nf_sum

    # ---- main arm: fp64 exact mode ----
    cfg = build_config(sim_config_kw={"turbulence": True}, precision="f64", integrator="rk45", seed=0,
                       env_id_offset=rank * n)
    env = bt.BatchedFixedWing(n, cfg=cfg, device=local)
    env.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    pool = [(torch.rand(n, 3, device=dev, generator=g) * 2 - 1).contiguous() for _ in range(16)]
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, nf = time_device_steps(env, pool, flush)
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / K
    value = world * n * K / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel: rk45_attempt_kernel, timed by itself ----
    # fw_set_profiling records CUDA events on the launch stream around each kernel of a step; the pass below repeats
    # the timed loop's conditions (fresh actions, L2 flush before every step) for up to 50 more steps.
    env.set_profiling(True)
    nf_prof = torch.zeros(2, dtype=torch.float64, device=dev)
    KP = max(1, min(K, 50))
    for k in range(KP):
        flush.add_(1.0)
        env.step(pool[k % len(pool)])
        nf_prof += env.get_field(bt.FIELD_NFEV).to(torch.float64).sum(0)
    prof = env.profile()
    env.set_profiling(False)
    nf_prof = (nf_prof / (n * KP)).cpu().numpy()
    peaks, peak_kind = measured_peaks()
    # numbers that only a profiler can give are READ from the committed capture summary, never typed in here
    ncu = {}
    ncu_path = os.path.join(ROOT, "profiles", "ncu_attempt_kernel.json")
    if os.path.exists(ncu_path):
        with open(ncu_path) as f:
            ncu = json.load(f)
    fp64_peak = bt.measure_fma_peak(local, "f64")
    fp32_peak = bt.measure_fma_peak(local, "f32")
    # algorithmic flops of one attempt-kernel launch: every RHS evaluation after the two of the init kernel, plus the
    # stage combinations / error norm of every attempt
    flops_attempt_env = (nf_prof[0] - 2.0) * (W_RHS + T_RHS) + nf_prof[1] * (W_ATT + T_ATT)
    achieved_tf = flops_attempt_env * n / (prof["integrate_ms"] * 1e-3) / 1e12
    flops_env_step = nf[0] * (W_RHS + T_RHS) + nf[1] * (W_ATT + T_ATT) + (W_ENV + T_ENV)
    step_tf = flops_env_step * n / (ms_per_step * 1e-3) / 1e12
    hbm_gbs = BYTES_PER_ENV_STEP_F64 * n / (ms_per_step * 1e-3) / 1e9
    ksum = prof["init_ms"] + prof["integrate_ms"] + prof["head_ms"]
    roofline = {"bound": "fp64", "kernel": "rk45_attempt_kernel<double, turbulence, 32>",
                "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                "kernel_ms": prof["integrate_ms"], "kernel_timing": "CUDA events around the kernel on its launch "
                "stream (fw_set_profiling), mean of %d launches, L2 flushed before each step" % KP,
                "peak_source": "DFMA micro-benchmark fw_measure_fma_peak on this GPU, same process (MEASURED_PEAKS.json "
                               "has no vector-pipe figure)",
                "flops_per_launch": flops_attempt_env * n, "flops_per_env_attempt_kernel": flops_attempt_env,
                "mean_rhs_evals": float(nf_prof[0]), "mean_rk_attempts": float(nf_prof[1]),
                "kernel_shares": {"rk45_init_kernel": prof["init_ms"] / ksum, "rk45_attempt_kernel": prof["integrate_ms"] / ksum,
                                  "head_kernel": prof["head_ms"] / ksum,
                                  "ms": [prof["init_ms"], prof["integrate_ms"], prof["head_ms"]]},
                "traffic": ncu.get("attempt_kernel_dram_bytes_per_launch"),
                "traffic_note": ncu.get("note", "no ncu capture committed (profiles/ncu_attempt_kernel.json absent)"),
                "executed_fp64_flop_per_env_step_ncu": ncu.get("executed_fp64_flop_per_env_step"),
                "whole_step": {"achieved": step_tf, "frac": step_tf / fp64_peak, "flops_per_env_step": flops_env_step,
                               "ms": ms_per_step},
                "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_gbs / peaks.get("hbm_gbs"), "peak_source": peak_kind + " (MEASURED_PEAKS.json)",
                        "bytes_per_env_step": BYTES_PER_ENV_STEP_F64},
                "fp32_peak_tflops": fp32_peak}
    env.close()
    del env

    # ---- e2e: VecEnv API, host numpy actions in, numpy obs/rew/done out ----
    e2e = None
    if not args.no_e2e:
        venv = FixedWingVecEnv(n, sim_config_kw={"turbulence": True}, device=local, seed=0, env_id_offset=rank * n,
                               precision="f64", integrator="rk45")
        venv.reset()
        rs = np.random.RandomState(rank)
        host_pool = venv.pinned_actions(8)          # the step's inputs sit in pinned host memory (bench contract)
        for a in host_pool:
            a[...] = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
        for w in range(W):
            venv.step(host_pool[w % 8])
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            obs, rew, done, infos = venv.step(host_pool[(W + k) % 8])
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n * K / dt, "unit": UNIT, "h2d_bytes_per_step": venv.h2d_bytes_per_step,
               "d2h_bytes_per_step": venv.d2h_bytes_per_step, "ms_per_step": dt / K * 1e3,
               "api": "FixedWingVecEnv.step(numpy float32 actions in pinned host memory) -> numpy obs, rewards, dones, "
                      "infos (lazy); H2D of the actions and one D2H of obs|rew|done|episode-end rows inside every step"}
        venv.close()
        del venv

    # ---- side measurements (not the headline): fp32 fixed-step mode ----
    extra = {}
    if not args.no_extra:
        for tag, prec, integ in (("fp32_rk4x4", "f32", "rk4"), ("fp64_rk4x4", "f64", "rk4")):
            c2 = build_config(sim_config_kw={"turbulence": True}, precision=prec, integrator=integ, rk4_substeps=4,
                              seed=0, env_id_offset=rank * n)
            e2 = bt.BatchedFixedWing(n, cfg=c2, device=local)
            e2.reset()
            ms2, _ = time_device_steps(e2, pool, flush)
            ms2 = max_over_ranks(ms2)
            extra[tag] = {"value": world * n * K / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K,
                          "note": "fixed-step mode, graded at its own tolerance (tests/test_gpu_env.py::test_fast_modes_at_their_stated_tolerance), not the "
                                  "parity path"}
            e2.close()

    # ---- BASELINE.json configs[1] ("C2"): 4096 envs, no turbulence, random actions — the pure dynamics step ----
    configs = {}
    if not args.no_extra:
        c2 = build_config(sim_config_kw={"turbulence": False}, precision="f64", integrator="rk45", seed=0,
                          env_id_offset=rank * 4096)
        e2 = bt.BatchedFixedWing(4096, cfg=c2, device=local)
        e2.reset()
        pool2 = [p_[:4096].contiguous() for p_ in pool]
        ms2, nf2 = time_device_steps(e2, pool2, flush, n_envs=4096)
        ms2 = max_over_ranks(ms2)
        configs["C2"] = {"workload": "4096 envs/GPU, turbulence off, U(-1,1)^3 actions, fp64 exact", "value": world * 4096 * K / (ms2 * 1e-3),
                         "unit": UNIT, "ms_per_step": ms2 / K, "mean_rhs_evals": float(nf2[0]), "mean_rk_attempts": float(nf2[1]),
                         "note": "4096 envs are 11 % of one wave of the persistent attempt kernel (37 888 lanes): the step "
                                 "lasts as long as its slowest env (max attempts x per-attempt latency), not as long as its work"}
        e2.close()
        # ---- C3 as written ("65536 envs ... sharded across 1/2/4/8"): STRONG scaling, 65536 / N envs per GPU ----
        if world > 1:
            ns = ENVS_PER_GPU // world
            c3s = build_config(sim_config_kw={"turbulence": True}, precision="f64", integrator="rk45", seed=0,
                               env_id_offset=rank * ns)
            e3 = bt.BatchedFixedWing(ns, cfg=c3s, device=local)
            e3.reset()
            pool3 = [p_[:ns].contiguous() for p_ in pool]
            ms3, _ = time_device_steps(e3, pool3, flush, n_envs=ns)
            ms3 = max_over_ranks(ms3)
            configs["C3_strong"] = {"workload": "65536 envs in total, %d per GPU" % ns, "value": world * ns * K / (ms3 * 1e-3),
                                    "unit": UNIT, "ms_per_step": ms3 / K, "scaling": "strong"}
            e3.close()
        else:
            configs["C3_strong"] = {"workload": "65536 envs in total = the headline line at n_gpus 1", "value": value,
                                    "unit": UNIT, "ms_per_step": ms_per_step, "scaling": "strong"}

    # ---- the same workload at 2x and 4x the batch: where the kernel leaves the tail-dominated regime ----
    if not args.no_extra:
        for mult in (2, 4):
            nb_ = ENVS_PER_GPU * mult
            cb_ = build_config(sim_config_kw={"turbulence": True}, precision="f64", integrator="rk45", seed=0,
                               env_id_offset=rank * nb_)
            eb_ = bt.BatchedFixedWing(nb_, cfg=cb_, device=local)
            eb_.reset()
            gb_ = torch.Generator(device=dev)
            gb_.manual_seed(2000 + rank)
            poolb = [(torch.rand(nb_, 3, device=dev, generator=gb_) * 2 - 1).contiguous() for _ in range(4)]
            Kb = min(K, 40)
            for w in range(5):
                eb_.step(poolb[w % 4])
            eb_.set_profiling(True)
            nfb = torch.zeros(2, dtype=torch.float64, device=dev)
            for k in range(Kb):
                flush.add_(1.0)
                eb_.step(poolb[k % 4])
                nfb += eb_.get_field(bt.FIELD_NFEV).to(torch.float64).sum(0)
            pb = eb_.profile()
            eb_.set_profiling(False)
            nfb = (nfb / (nb_ * Kb)).cpu().numpy()
            msb = max_over_ranks(pb["init_ms"] + pb["integrate_ms"] + pb["head_ms"])
            tf = ((nfb[0] - 2.0) * (W_RHS + T_RHS) + nfb[1] * (W_ATT + T_ATT)) * nb_ / (pb["integrate_ms"] * 1e-3) / 1e12
            configs["C3_batch_x%d" % mult] = {"workload": "C3 with %d envs per GPU" % nb_, "value": world * nb_ / (msb * 1e-3), "unit": UNIT,
                                             "ms_per_step": msb, "attempt_kernel_frac": tf / fp64_peak,
                                             "note": "sum of the three kernels' CUDA-event times per step (profiling mode); the "
                                                     "tail of stragglers (9 sequential attempts) weighs less the larger the batch"}
            eb_.close()
            del eb_, poolb

    # ---- the HBM-bound kernels of the PPO path at C4 sizes: GAE and the per-step rollout glue ----
    hbm_kernels = {}
    if not args.no_extra:
        hbm_peak = peaks.get("hbm_gbs")
        Tg, Ng = 2048, 8192                                   # SB3's default n_steps at 8192 envs/GPU: 335 MB per pass
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        rew_g, val_g = (torch.randn(Tg, Ng, device=dev, generator=gen) for _ in range(2))
        done_g = (torch.rand(Tg, Ng, device=dev, generator=gen) < 0.0005).float()
        lv_g, ld_g = torch.randn(Ng, device=dev, generator=gen), torch.zeros(Ng, dtype=torch.uint8, device=dev)
        for _ in range(3):
            bt.gae(rew_g, val_g, done_g, lv_g, ld_g)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a_, b_ in evs:                                    # 335 MB of inputs + outputs per pass: larger than L2
            a_.record(); bt.gae(rew_g, val_g, done_g, lv_g, ld_g); b_.record()
        torch.cuda.synchronize()
        ms_g = float(np.median([a_.elapsed_time(b_) for a_, b_ in evs]))
        gbs = 20.0 * Tg * Ng / (ms_g * 1e-3) / 1e9
        hbm_kernels["gae"] = {"kernel": "gae_stream_kernel (fw_gae)", "T": Tg, "N": Ng, "bytes_per_transition": 20,
                              "ms": ms_g, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                              "timing": "median of 10 launches, CUDA events; working set 335 MB > L2"}
        del rew_g, val_g, done_g
        # fw_rollout_post_step: VecNormalize + RunningMeanStd + RolloutBuffer.add for one step of 8192 envs
        from tum_adlr_deep_reinforcement_learning_b200.buffers import (DeviceVecNormalize, RolloutBuffer, fused_post_step,
                                                                        rollout_scratch_doubles)
        nb, od = 8192, 14
        norm_b = DeviceVecNormalize(nb, obs_dim=od, device=dev)
        buf_b = RolloutBuffer(64, nb, obs_dim=od, device=dev)
        ob_b, rw_b = torch.randn(nb, od, device=dev), torch.randn(nb, device=dev)
        dn_b = torch.zeros(nb, dtype=torch.uint8, device=dev)
        ac_b, vl_b, lp_b = torch.randn(nb, 3, device=dev), torch.randn(nb, device=dev), torch.randn(nb, device=dev)
        lo_b, ld_b = torch.zeros(nb, od, device=dev), torch.zeros(nb, device=dev)
        rr_b, rl_b = torch.zeros(nb, dtype=torch.float64, device=dev), torch.zeros(nb, dtype=torch.float64, device=dev)
        es_b, sc_b = torch.zeros(3, dtype=torch.float64, device=dev), torch.zeros(rollout_scratch_doubles(od), dtype=torch.float64, device=dev)
        times = []
        for it in range(40):
            buf_b.pos = it % 64
            flush.add_(1.0)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            fused_post_step(norm_b, buf_b, ob_b, rw_b, dn_b, ac_b, vl_b, lp_b, lo_b, ld_b, rr_b, rl_b, es_b, sc_b)
            b_.record()
            torch.cuda.synchronize()
            if it >= 8:
                times.append(a_.elapsed_time(b_))
        ms_b = float(np.median(times))
        bytes_env = 4 * (od + 1 + 3 + 1 + 1) + 1 + 2 * 4 * od + 2 * 4 + 2 * 8 * 3 + 4 * (od + 3 + 4)   # in, state r+w, row out
        gbs = bytes_env * nb / (ms_b * 1e-3) / 1e9
        hbm_kernels["rollout_post_step"] = {"kernel": "rollout_stats/moments/apply kernels (fw_rollout_post_step)", "N": nb,
                                            "bytes_per_env": bytes_env, "ms": ms_b, "achieved": gbs, "peak": hbm_peak,
                                            "unit": "GB/s", "frac": gbs / hbm_peak,
                                            "note": "2.7 MB per call = 0.4 us at the HBM peak: three dependent launches, bound by "
                                                    "launch latency, not bandwidth; inside ppo.PPO they are replayed from a CUDA graph"}

    # ---- the SAC replay ring (config C5 sizes: 1e6 rows, insert 1024, sample 4096) ----
    if not args.no_extra:
        from tum_adlr_deep_reinforcement_learning_b200.sac import ReplayBuffer
        rbuf = ReplayBuffer(1_000_000, device=dev, seed=1)
        rbuf.rows.normal_()
        rbuf.size_dev.fill_(1_000_000); rbuf.pos, rbuf.full = 0, True
        nrm = DeviceVecNormalize(1, obs_dim=14, device=dev)
        o_i, a_i, r_i = torch.randn(1024, 14, device=dev), torch.randn(1024, 3, device=dev), torch.randn(1024, device=dev)
        d_i = torch.zeros(1024, dtype=torch.uint8, device=dev)
        t_ins, t_smp = [], []
        for it in range(40):
            flush.add_(1.0)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(); rbuf.add(o_i, o_i, a_i, r_i, d_i); e1.record(); rbuf.sample(4096, norm=nrm); e2.record()
            torch.cuda.synchronize()
            if it >= 8:
                t_ins.append(e0.elapsed_time(e1)); t_smp.append(e1.elapsed_time(e2))
        rowb = rbuf.row_floats * 4
        for tag, ms_k, nbytes, what in (("replay_insert", float(np.median(t_ins)), 1024 * (rowb + 4 * 33), "1024 rows of %d B written + their sources read" % rowb),
                                        ("replay_sample", float(np.median(t_smp)), 4096 * (rowb + 4 * 33), "4096 random rows of %d B read + the batch written" % rowb)):
            gbs = nbytes / (ms_k * 1e-3) / 1e9
            hbm_kernels[tag] = {"kernel": "fw_" + tag, "ms": ms_k, "bytes": nbytes, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": gbs / hbm_peak, "note": what + "; two launches, bound by launch latency at these sizes"}
        del rbuf

    # ---- PPO train env-steps/s (BASELINE.json second metric, config C4: 8192 envs/GPU, updates included) ----
    ppo = None
    if not args.no_extra:
        from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
        n_ppo, n_steps, iters = 8192, 32, 10
        venv = FixedWingVecEnv(n_ppo, sim_config_kw={"turbulence": True}, device=local, seed=0, env_id_offset=rank * n_ppo)
        algo = PPO(venv, n_steps=n_steps, batch_size=n_ppo * n_steps // 4, n_epochs=10, ent_coef=0.01,
                   dist=dist if world > 1 else None)
        algo.learn(total_timesteps=3 * world * n_ppo * n_steps)        # warm-up: eager pass, CUDA-graph captures
        assert algo._rollout_graph is not None and algo._train_graph, "PPO must run from captured CUDA graphs in the bench"
        barrier()
        t0 = time.perf_counter()
        algo.learn(total_timesteps=algo.num_timesteps + iters * world * n_ppo * n_steps)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        ppo = {"value": iters * world * n_ppo * n_steps / dt, "unit": "env-steps/s (rollout + GAE + 10-epoch update)",
               "envs_per_gpu": n_ppo, "n_steps": n_steps, "minibatch": n_ppo * n_steps // 4, "n_epochs": 10,
               "hyper_parameters": "4 minibatches x 10 epochs per rollout of 32 steps (nminibatches = 4 is the PPO2 default of the "
                                   "reference's training script); reaches the reference's 1 M-step reward (-405) in 16.6 s of "
                                   "wall-clock, results/r02_ppo_curve_60s_4minibatches.json (8 minibatches: 16.2 s, 7.6e6 env-steps/s)",
               "iterations_timed": iters, "ep_rew_mean": algo.logs[-1]["ep_rew_mean"], "cuda_graphs": True,
               "note": "policy 2x64 tanh MLP in PyTorch (split-K weight gradients), loss block fused in fw_ppo_loss, rollout "
                       "and minibatch update replayed as CUDA graphs (value branch on a second stream); one gradient all-reduce per optimiser step when "
                       "n_gpus > 1 (NCCL, captured in the update graph); learning curves in results/"}
        venv.close()

    # ---- SAC (BASELINE.json config C5: 1024 envs, replay ring in HBM, turbulence on), N = 1 only ----
    sac = None
    if not args.no_extra and world == 1:
        from tum_adlr_deep_reinforcement_learning_b200.sac import SAC
        venv = FixedWingVecEnv(1024, sim_config_kw={"turbulence": True}, device=local, seed=0)
        salgo = SAC(venv, buffer_size=1_000_000, batch_size=4096, gradient_steps=2, learning_starts=10_000)
        salgo.learn(total_timesteps=40 * 1024)                          # warm-up: random phase, graph captures
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n0 = salgo.num_timesteps
        salgo.learn(total_timesteps=n0 + 1500 * 1024, log_every=10 ** 9)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert salgo._env_graph and salgo._train_graph, "SAC must run from captured CUDA graphs in the bench"
        sac = {"value": (salgo.num_timesteps - n0) / dt, "unit": "env-steps/s (1 env step of 1024 envs + 2 gradient steps of batch 4096)",
               "gradient_steps_per_s": 2 * (salgo.num_timesteps - n0) / 1024 / dt, "envs": 1024, "replay_capacity": 1_000_000,
               "cuda_graphs": True,
               "note": "env step and gradient step replayed as CUDA graphs; networks 2x256 ReLU in PyTorch"}
        venv.close()

    # ---- CPU baseline (rank 0 only, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        rate, sample = cpu_port_rate(cfg, args.cpu_budget_s, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "python_reference": python_reference_rate(args.cpu_budget_s)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(n), "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 4 * K, "gpu_launches_note": "rk45_init_kernel + rk45_attempt_kernel + head_kernel per step on the launch stream, refill_kernel on the side stream",
                "clocks": clocks, "modes": extra, "configs": configs, "hbm_kernels": hbm_kernels, "ppo": ppo, "sac": sac}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that contain NCCL kernels (the PPO update) are alive until interpreter shutdown, and
        # destroy_process_group was seen to wait forever behind them: everything is flushed and synchronised, so leave
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
