"""CPU, world_size 2, gloo: host-side logic of the molecule-sharded data-parallel path (molkgnn_b200/dp.py).

The CUDA kernels cannot run here, so each rank produces the gradients of its shard with the CPU oracle and the test
checks what the product's DP layer is responsible for: shard bounds, which parameters enter the flat bucket, and that
the all-reduced (summed) shard gradients equal the gradients of the unsharded batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L1 = (3, 4, 5, 6)
LN = (2, 3, 4, 5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads_via_oracle(net, mols, wscale=1.0):
    """fills p.grad of the (CPU) drop-in module from the oracle's autograd on the given molecules"""
    from molkgnn_b200 import synth
    from oracle import molkgnn_oracle as orc
    from tests.helpers import params_from_module
    b = synth.collate(mols)
    params = params_from_module(net, requires_grad=True)
    N = b["x"].shape[0]
    bk = orc.buckets_to_torch(orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"]))
    h = orc.molgcn_forward(params, torch.from_numpy(b["x"]), torch.from_numpy(b["edge_index"]), bk)
    (h.sum() * wscale).backward()
    for li, layer in enumerate(net.layers):
        for d, kc in enumerate(layer.trainable_kernelconv_set):
            for n, t in params[li][d].items():
                getattr(kc, n).grad = None if t.grad is None else t.grad.clone()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from molkgnn_b200.dp import GradBucket, shard_bounds
    mols = synth.make_molecules(10, seed=5)
    bounds = shard_bounds([m.num_nodes for m in mols], world)
    torch.manual_seed(0)
    net = mk.MolGCN(2, *L1, *LN, x_dim=28, p_dim=3, edge_attr_dim=7)
    lo, hi = bounds[rank]
    _grads_via_oracle(net, mols[lo:hi])
    bucket = GradBucket(net, world, average=False)
    flat = bucket.allreduce()
    if rank == 0:
        torch.save({"flat": flat, "numel": bucket.numel, "bounds": bounds,
                    "names": [n for n, p in net.named_parameters() if any(p is q for q in bucket.params)]}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_shard_grads_equals_full_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    sys.path.insert(0, ROOT)
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from molkgnn_b200.dp import GradBucket
    mols = synth.make_molecules(10, seed=5)
    torch.manual_seed(0)
    net = mk.MolGCN(2, *L1, *LN, x_dim=28, p_dim=3, edge_attr_dim=7)
    _grads_via_oracle(net, mols)
    ref = torch.cat([p.grad.reshape(-1) for p in GradBucket(net, 1).params])
    assert res["flat"].shape == ref.shape
    assert torch.allclose(res["flat"], ref, rtol=1e-4, atol=1e-6)
    # bucket membership: exactly the parameters the reference differentiates
    assert not any(n.endswith(("p_support", "length_sc_weight", "angle_sc_weight")) for n in res["names"])
    assert len(res["names"]) == 2 * 4 * 6
    (a0, a1), (b0, b1) = res["bounds"]
    assert a0 == 0 and a1 == b0 and b1 == 10 and 0 < a1 < 10


def test_shard_bounds_balance_by_atoms():
    from molkgnn_b200.dp import shard_bounds
    sizes = [30, 10, 10, 10, 30, 10]
    b = shard_bounds(sizes, 2)
    assert b == [(0, 3), (3, 6)]
    b8 = shard_bounds([25] * 64, 8)
    assert [hi - lo for lo, hi in b8] == [8] * 8
    assert shard_bounds([5, 5], 4)[-1][1] == 2
