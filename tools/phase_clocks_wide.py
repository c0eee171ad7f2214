"""In-kernel phase clocks of the wide kernels (profiling build, -DMK_PHASE_CLOCKS), configs[2] model.

    python -m molkgnn_b200.build --phase-clocks          # here (cross-compile), then on the GPU box:
    MOLKGNN_B200_LIB=molkgnn_b200/libmolkgnn_b200_prof.so python tools/phase_clocks_wide.py [molecules] [steps]
Cycles are per CTA and step (sum over the CTAs / number of CTAs / steps); thread 0 of the consumers, the ring lane, MMA lane 0.
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
b = synth.make_batch(B, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
Lk = (40, 80, 120, 200)
net = mk.MolGCN(5, *Lk, *Lk, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 440, device=dev)
L = _lib.lib()


def step():
    x = t["x"].detach().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    h.backward(wout)
    net.zero_grad(set_to_none=True)


def read(name, n):
    f = getattr(L, name)
    f.restype = C.c_int
    a = (C.c_ulonglong * n)()
    assert f(a) == 0
    return [int(v) for v in a]


for _ in range(2):
    step()
torch.cuda.synchronize()
read("molkgnn_debug_phase_clocks_wfwd", 48)
read("molkgnn_debug_phase_clocks_wbwd", 96)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ncta = L.molkgnn_num_sms()
f = read("molkgnn_debug_phase_clocks_wfwd", 48)
w = read("molkgnn_debug_phase_clocks_wbwd", 96)
k = 1.0 / (ncta * steps)
out = {"molecules": B, "steps": steps, "ms_per_step": e0.elapsed_time(e1) / steps, "cycles": "per CTA and step (5 layers)",
       "fwd_consumer": {n: round(v * k) for n, v in zip(
           ["prologue", "meta/eh wait + dup flags", "wait MMA", "dump", "pairs", "sync after pairs"], f[:6])},
       "fwd_ring_blockA": {"wait stage free": round(f[17] * k), "other (issue)": round(f[16] * k)},
       "fwd_mma0": {"wait accumulator drained": round(f[33] * k), "wait node stage": round(f[34] * k), "wait block stage": round(f[35] * k),
                    "issue": round(f[36] * k), "other": round(f[32] * k)}}
for ph, nm in ((0, "bwd_X"), (1, "bwd_G")):
    c = w[48 * ph:48 * ph + 48]
    out[nm + "_consumer"] = {n: round(v * k) for n, v in zip(
        ["prologue", "wait MMAs of unit u-2 / copies / tail", "Wt clear", "rank-0 scatter", "chains", "fence+sync", "flush", "wait chunk's last MMAs + final"], c[:8])}
    out[nm + "_ring"] = {"wait free stage": round(c[17] * k), "other": round(c[16] * k)}
    out[nm + "_mma0"] = {"wait Wt": round(c[33] * k), "wait stage": round(c[34] * k), "issue": round(c[35] * k), "other": round(c[32] * k)}
print(json.dumps(out))
