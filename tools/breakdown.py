"""Per-kernel-group CUDA-event breakdown of the bench step (library profiler).  python tools/breakdown.py [molecules] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth, functional as Fn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
b = synth.make_batch(B, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 110, device=dev)


def step():
    x = t["x"].detach().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    h.backward(wout)
    net.zero_grad(set_to_none=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
Fn.profile_start()
for _ in range(steps):
    step()
bd = Fn.profile_stop()
out = {"ms_per_step": ms, "us_per_step": {k: round(1e3 * v[1] / steps, 1) for k, v in bd.items()},
       "launch_groups_per_step": {k: v[0] / steps for k, v in bd.items()}}
out["sum_us"] = round(sum(out["us_per_step"].values()), 1)
print(json.dumps(out, indent=1))
