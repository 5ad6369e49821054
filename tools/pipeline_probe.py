"""Open-loop (random-action) workload with the batch split into G independent groups on G streams: does the tail of one
group's attempt kernel overlap the init / head / attempt start of the others?  Prototype through G separate handles
(global env ids keep the results those of one handle)."""
import sys, json; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = 120
res = {}
for G in (1, 2, 4, 8, 16):
    n = N // G
    envs = [bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True}, env_id_offset=g * n)) for g in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    for e in envs: e.reset()
    torch.cuda.synchronize()
    def run(k):
        main = torch.cuda.current_stream()
        for s in streams: s.wait_stream(main)
        for _ in range(k):
            for e, s in zip(envs, streams):
                with torch.cuda.stream(s): e.step_random(1)
        for s in streams: main.wait_stream(s)
    run(30)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    # the same through one CUDA graph (launch overhead off the table)
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    gms = None
    try:
        with torch.cuda.stream(cap):
            for e in envs: e.join()
            torch.cuda.synchronize()
            g.capture_begin()
            run(20)
            for e in envs: e.join()
            g.capture_end()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / 100
    except Exception as ex:
        print("graph capture failed:", repr(ex)[:200])
    res[G] = dict(ms=ms, rate=N / ms * 1e3, graph_ms=gms, graph_rate=(N / gms * 1e3 if gms else None))
    print("N=%d G=%2d  ms per N-env step %.4f  %.3e env-steps/s | graph %s" % (N, G, ms, N / ms * 1e3,
          ("%.4f ms %.3e" % (gms, N / gms * 1e3)) if gms else "-"), flush=True)
    for e in envs: e.close()
    del envs, g
json.dump(res, open("gpurun_out/pipeline_probe_%d.json" % N, "w"), indent=1)
