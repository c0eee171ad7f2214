"""One-shot NVLink all-reduce (csrc/oneshot.cu) against NCCL, under torchrun on >= 2 GPUs of one node:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/oneshot_check.py
Checks: results bitwise identical on all ranks, equal to NCCL's sum within fp32 rounding, over many back-to-back steps with
skewed ranks (slot / flag reuse); then times both on the launching stream.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from molkgnn_b200.dp import OneShotAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 121872 + 3                                   # not a multiple of 4 on purpose
osr = OneShotAllReduce.create(N + 64)
out = {"world": world, "created": osr is not None}
if osr is None:
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0)

g = torch.Generator(device=dev).manual_seed(1234 + rank)
steps = 64
inputs = [torch.randn(N, device=dev, generator=g) for _ in range(8)]
max_rel, all_equal = 0.0, True
mine_list, ref_list = [], []
for s in range(steps):
    x = inputs[s % 8] * (1.0 + s)
    ref = x.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    mine = x.clone()
    if (s + rank) % 3 == 0:
        torch.cuda._sleep(2_000_000)             # skew the ranks (~1 ms) without a host sync in between
    osr.allreduce(mine, average=False)
    mine_list.append(mine)
    ref_list.append(ref)
torch.cuda.synchronize()
osr.check()
for mine, ref in zip(mine_list, ref_list):
    max_rel = max(max_rel, float((mine - ref).abs().max() / ref.abs().max()))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    all_equal = all_equal and all(torch.equal(gathered[0], t) for t in gathered[1:])
# mean
x = inputs[0].clone()
ref = x.clone()
dist.all_reduce(ref, op=dist.ReduceOp.AVG)
osr.allreduce(x, average=True)
torch.cuda.synchronize()
out["avg_rel"] = float((x - ref).abs().max() / ref.abs().max())
out["max_rel_vs_nccl"] = max_rel
out["bitwise_identical_across_ranks"] = bool(all_equal)


def timed(fn, iters=200):
    buf = inputs[1].clone()
    for _ in range(10):
        fn(buf)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out["us_per_call_nccl_avg"] = timed(lambda b: dist.all_reduce(b, op=dist.ReduceOp.AVG))
out["us_per_call_oneshot"] = timed(lambda b: osr.allreduce(b, average=True))
osr.check()
out["floats"] = N
if rank == 0:
    print(json.dumps(out))
osr.close()
dist.destroy_process_group()
