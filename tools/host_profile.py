"""cProfile of the host side of the training step (plan staged ahead, as in bench.py).  python tools/host_profile.py"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402
from molkgnn_b200.data import DevicePrefetcher  # noqa: E402

dev = torch.device("cuda", 0)
b = synth.make_batch(4096, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 110, device=dev)
pf = DevicePrefetcher(dev)


def run(k):
    nxt = pf.put(device_batch=t, build_plan=True)
    for i in range(k):
        tt, plan = pf.get(nxt)
        if i + 1 < k:
            nxt = pf.put(device_batch=t, build_plan=True)
        x = tt["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=tt["edge_index"], edge_attr=tt["edge_attr"], p=tt["p"], save_score=False, plan=plan)
        h.backward(wout)
        net.zero_grad(set_to_none=True)


run(10)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
run(200)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
