"""Per-step timeline of the end-to-end loop (bench.py e2e_run) for a SHORT run: where a fixed start-up / drain cost sits.
    python tools/e2e_trace.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402
from molkgnn_b200.data import DevicePrefetcher  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
b = synth.make_batch(4096, seed=0)
host = {k: torch.from_numpy(b[k]).pin_memory() for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(host["x"].shape[0], 110, device=dev)
pf = DevicePrefetcher(dev)
params = list(net.parameters())
loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]


def run(k, trace=None):
    q = [pf.put(host_batch=host, build_plan=True) for _ in range(min(pf.depth, k))]
    pending = None
    for i in range(k):
        t0 = time.perf_counter()
        t, plan = pf.get(q.pop(0))
        t1 = time.perf_counter()
        seg0 = torch.cuda.memory_stats()["segment.all.allocated"]
        if i + pf.depth < k:
            q.append(pf.put(host_batch=host, build_plan=True))
        t2 = time.perf_counter()
        seg1 = torch.cuda.memory_stats()["segment.all.allocated"]
        x = t["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
        h.backward(wout)
        loss = (h.detach() * wout).sum()
        buf = loss_host[i & 1]
        buf.copy_(loss, non_blocking=True)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        for p in params:
            p.grad = None
        t3 = time.perf_counter()
        if pending is not None:
            pending[1].synchronize()
            float(pending[0])
        t4 = time.perf_counter()
        pending = (buf, ev)
        if trace is not None:
            trace.append((ev, t0, t1, t2, t3, t4, seg1 - seg0, torch.cuda.memory_stats()["segment.all.allocated"] - seg1))
    pending[1].synchronize()


run(3)
run(5)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True)
tw0 = time.perf_counter()
e0.record()
tr = []
run(K, tr)
e1 = torch.cuda.Event(enable_timing=True)
e1.record()
torch.cuda.synchronize()
print(f"total {e0.elapsed_time(e1):.3f} ms for {K} steps = {e0.elapsed_time(e1) / K:.3f} ms per step")
prev = e0
for i, (ev, t0, t1, t2, t3, t4, dseg_put, dseg_step) in enumerate(tr):
    print(f"step {i:2d}: gpu +{prev.elapsed_time(ev):6.3f} ms | host get {1e3 * (t1 - t0):6.3f} put {1e3 * (t2 - t1):6.3f} "
          f"fwd+bwd {1e3 * (t3 - t2):6.3f} wait {1e3 * (t4 - t3):6.3f} | host t={1e3 * (t0 - tw0):7.3f} | cudaMalloc segments: put {dseg_put} step {dseg_step}")
    prev = ev
