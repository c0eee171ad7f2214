"""Profiling driver: a few steps of the bench workload (for ncu).   python tools/prof_step.py [molecules] [steps] [wide]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
b = synth.make_batch(B, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
WIDE = len(sys.argv) > 3 and sys.argv[3] == "wide"          # BASELINE configs[2]
Lk = (40, 80, 120, 200) if WIDE else (10, 20, 30, 50)
net = mk.MolGCN(5 if WIDE else 3, *Lk, *Lk, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], sum(Lk), device=dev)
for _ in range(steps):
    x = t["x"].detach().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    h.backward(wout)
    net.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done", t["x"].shape)
