"""cProfile of the host side of the training step, as bench.py drives it (batches staged two steps ahead).

    python tools/host_profile.py [host|store|device]     # where the batch comes from (default host: pinned-memory copies)
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402
from molkgnn_b200.data import DevicePrefetcher  # noqa: E402
from molkgnn_b200.store import MoleculeStore  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "host"
dev = torch.device("cuda", 0)
B = 4096
b = synth.make_batch(B, seed=0)
host = {k: torch.from_numpy(b[k]).pin_memory() for k in ("x", "p", "edge_index", "edge_attr")}
t = {k: v.to(dev) for k, v in host.items()}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 110, device=dev)
pf = DevicePrefetcher(dev)
params = list(net.parameters())
ptr_h = b["ptr"]
eptr = np.searchsorted(b["edge_index"][0], ptr_h)
local_ei = b["edge_index"] - np.repeat(ptr_h[:-1], np.diff(eptr))[None, :]
store = MoleculeStore(t["x"], t["p"], torch.from_numpy(local_ei).to(dev), t["edge_attr"], torch.from_numpy(ptr_h).to(dev),
                      torch.from_numpy(eptr).to(dev))
ids = [torch.randperm(B).pin_memory() for _ in range(4)]
loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]


def put(i):
    if mode == "host":
        return pf.put(host_batch=host, build_plan=True)
    if mode == "store":
        return pf.put(store_batch=(store, ids[i & 3]), build_plan=True)
    return pf.put(device_batch=t, build_plan=True)


def run(k):
    q = [put(j) for j in range(min(pf.depth, k))]
    pending = None
    for i in range(k):
        tt, plan = pf.get(q.pop(0))
        if i + pf.depth < k:
            q.append(put(i + pf.depth))
        x = tt["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=tt["edge_index"], edge_attr=tt["edge_attr"], p=tt["p"], save_score=False, plan=plan)
        h.backward(wout)
        loss = (h.detach() * wout).sum()
        buf = loss_host[i & 1]
        buf.copy_(loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        for p in params:
            p.grad = None
        if pending is not None:
            pending[1].synchronize()
            float(pending[0])
        pending = (buf, ev)
    torch.cuda.synchronize()


run(10)
t0 = time.perf_counter()
run(200)
print(f"mode {mode}: wall {1e3 * (time.perf_counter() - t0) / 200:.3f} ms per step (no profiler)")
pr = cProfile.Profile()
pr.enable()
run(200)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
