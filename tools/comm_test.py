"""fw_comm_allreduce_adam against ncclAllReduce + divide + fw_adam_clip_step on N GPUs of one node:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_test.py
Checks bit-equality of the parameters after every step (the peer kernel adds the ranks in rank order; for 2 ranks that is
NCCL's sum as well), replays both from CUDA graphs and prints the time per optimiser step of each."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_adlr_deep_reinforcement_learning_b200.ppo import ActorCritic, FlatAdam  # noqa: E402

os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def make():
    torch.manual_seed(0)
    m = ActorCritic().to(dev)
    return m, FlatAdam(m, lr=3e-4, eps=1e-5, max_grad_norm=0.5)


ma, oa = make()          # NCCL path
mb, ob = make()          # peer-memory path
assert ob.enable_peer_allreduce(dist), "peer buffers could not be mapped"
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for step in range(30):
    grad = torch.randn(oa.grad.numel(), device=dev, generator=g) * (0.01 + 0.2 * (step % 3))
    oa.grad.copy_(grad)
    ob.grad.copy_(grad)
    dist.all_reduce(oa.grad)
    oa.grad.div_(world)
    oa.step()
    ob.step()
    worst = max(worst, float((oa.flat - ob.flat).abs().max()))
    assert torch.equal(oa.grad, ob.grad) or world > 2, step
ref = oa.flat.clone()
dist.all_reduce(ref, op=dist.ReduceOp.MAX)
assert torch.equal(ref, oa.flat), "replicas drifted"
tol = 0.0 if world == 2 else 1e-6
assert worst <= tol, worst
assert not ob.comm_error()


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            fn()
    gr.replay(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters // 20):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def nccl_step():
    dist.all_reduce(oa.grad)
    oa.grad.div_(world)
    oa.step()


t_nccl = timed(nccl_step)
t_peer = timed(ob.step)
assert not ob.comm_error()
if rank == 0:
    print(json.dumps({"world": world, "max_param_diff": worst, "us_per_step_nccl_allreduce_div_adam": t_nccl,
                      "us_per_step_peer_allreduce_adam": t_peer, "params": oa.flat.numel()}), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
