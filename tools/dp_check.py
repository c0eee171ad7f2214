#!/usr/bin/env python
"""Multi-GPU parity of the data-parallel step on hardware: every rank runs the CUDA forward/backward on
ITS shard of a batch (molecules sharded by atoms, dp.shard_bounds), GradBucket.allreduce(average=False) sums the flat
kernel-gradient buffer over NCCL (or the one-shot NVLink kernel), and the result must equal the gradients of the FULL batch
computed by one rank alone -- for both exchange paths.  Launch under torchrun; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_check(molecules=256):
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from molkgnn_b200.dp import GradBucket, shard_bounds
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    mols = synth.make_molecules(molecules, seed=77)
    bounds = shard_bounds([m.num_nodes for m in mols], world)

    def grads_of(sub, oneshot, reduce):
        torch.manual_seed(0)
        net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
        b = synth.collate(sub)
        t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
        x = t["x"].requires_grad_(True)
        h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
        h.sum().backward()
        bucket = GradBucket(net, world, average=False, oneshot=oneshot)
        if reduce:
            bucket.allreduce()
            bucket.check()
        torch.cuda.synchronize()
        g = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        return g, bucket.oneshot is not None

    full, _ = grads_of(mols, False, False)                      # every rank: the unsharded batch, no exchange
    out = {"world": world, "molecules": molecules, "bounds": bounds}
    lo, hi = bounds[rank]
    for name, oneshot in (("nccl", False), ("oneshot", True)):
        got, used = grads_of(mols[lo:hi], oneshot, True)
        worst = 0.0
        for n, ref in full.items():
            if n.endswith("_sc_weight"):
                continue
            worst = max(worst, float((got[n] - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
        w = torch.tensor([worst], device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        out[name] = {"max_rel_err_vs_full_batch": float(w.item()), "oneshot_active": bool(used)}
    return out


if __name__ == "__main__":
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_check()
    if dist.get_rank() == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
