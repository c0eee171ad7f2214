"""Kernel timeline of steady-state steps (CUPTI through torch.profiler): start / duration of every kernel, idle gaps on the GPU.
    python tools/gap_trace.py [molecules] [steps]  ->  one line per kernel of the last step + a summary (stdout)"""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402
from molkgnn_b200.data import DevicePrefetcher  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
b = synth.make_batch(B, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 110, device=dev)
pf = DevicePrefetcher(dev)
params = list(net.parameters())


def run(k):
    nxt = pf.put(device_batch=t, build_plan=True)
    for i in range(k):
        tt, plan = pf.get(nxt)
        if i + 1 < k:
            nxt = pf.put(device_batch=t, build_plan=True)
        x = tt["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=tt["edge_index"], edge_attr=tt["edge_attr"], p=tt["p"], save_score=False, plan=plan)
        h.backward(wout)
        for p in params:
            p.grad = None


run(20)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(steps)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
# union of busy intervals -> idle time
busy, end = 0.0, None
cur_s, cur_e = None, None
for e in evs:
    s, f = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, f
    else:
        cur_e = max(cur_e, f)
busy += cur_e - cur_s
span = evs[-1].time_range.end - t0
print(f"span {span:.1f} us for {steps} steps = {span / steps:.1f} us/step, GPU busy (union of kernels) {busy / steps:.1f} us/step, "
      f"idle {(span - busy) / steps:.1f} us/step")
# last full step: from the last k_x_images on
names = [e.name for e in evs]
starts = [i for i, n in enumerate(names) if "k_x_images" in n]
i0 = starts[-2] if len(starts) > 1 else 0
i1 = starts[-1]
prev_end = None
for e in evs[i0:i1]:
    s, f = e.time_range.start - t0, e.time_range.end - t0
    gap = "" if prev_end is None else f"{s - prev_end:7.1f}"
    print(f"{s:10.1f} {f - s:8.1f} gap_prev_end {gap:>8s}  {e.name[:60]}")
    prev_end = f if prev_end is None else max(prev_end, f)
