"""In-kernel phase clocks of the tile kernels (profiling build, -DMK_PHASE_CLOCKS) + host issue time per step.

    python -m molkgnn_b200.build --phase-clocks          # here (cross-compile), then on the GPU box:
    MOLKGNN_B200_LIB=molkgnn_b200/libmolkgnn_b200_prof.so python tools/phase_clocks.py [molecules] [steps]
"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import molkgnn_b200 as mk  # noqa: E402
from molkgnn_b200 import synth, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
b = synth.make_batch(B, seed=0)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
torch.manual_seed(0)
net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
wout = torch.randn(t["x"].shape[0], 110, device=dev)
L = _lib.lib()
has_clk = hasattr(L, "molkgnn_debug_phase_clocks_bwd")
has_sf = hasattr(L, "molkgnn_debug_phase_clocks_sfwd")


def step(plan=None):
    x = t["x"].detach().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
    h.backward(wout)
    net.zero_grad(set_to_none=True)


def read(name, n):
    f = getattr(L, name)
    f.restype = C.c_int
    a = (C.c_ulonglong * n)()
    assert f(a) == 0
    return [int(v) for v in a]


for _ in range(3):
    step()
torch.cuda.synchronize()
if has_clk:
    read("molkgnn_debug_phase_clocks_bwd", 16)
    read("molkgnn_debug_phase_clocks_fwd", 32)
    read("molkgnn_debug_phase_clocks_coef", 16)
    if has_sf:
        read("molkgnn_debug_phase_clocks_sfwd", 48)
    if hasattr(L, "molkgnn_debug_phase_clocks_bwdp"):
        read("molkgnn_debug_phase_clocks_bwdp", 48)
out = {}
# (1) full step, plan rebuilt every step (one stream sync inside the bucket pass)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
out["ms_per_step_plan_rebuilt"] = e0.elapsed_time(e1) / steps
if has_clk:
    bw = read("molkgnn_debug_phase_clocks_bwd", 16)
    fw = read("molkgnn_debug_phase_clocks_fwd", 32)
    cf = read("molkgnn_debug_phase_clocks_coef", 16)
    names_c = ["prologue", "barrier + issue next", "wait data", "pairs", "pairs barrier", "bond sums", "final"]
    ncta = 148
    out["coef_tile_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_c, cf)}
    names_b = ["prologue", "wait tile copy", "(unused)", "ranks 1-3", "barrier before MMA", "img wait + MMA issue",
               "MMA wait", "Wt clear", "dxh epilogue", "G store", "rank-0 scatter block 0", "rank-0 scatter block 1",
               "rank-0 scatter block 2", "rank-0 scatter block 3", "set-up in front of the scatter"]
    names_fc = ["block set-up", "wait copy", "wait MMA", "dump", "dupflags+sync", "epilogue", "sync", "teardown"]
    names_fp = ["block set-up", "wait bfree", "wait img buffer", "copy wait", "wait tfree", "MMA issue"]
    out["bwd_tile_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_b, bw)}
    out["fwd_tile_consumer_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_fc, fw[:16])}
    out["fwd_tile_producer_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_fp, fw[16:])}
    if has_sf:
        sf = read("molkgnn_debug_phase_clocks_sfwd", 48)
        names_sc = ["prologue", "wait meta/bonds/x", "layer-0 image + S0", "wait MMA", "dump + S1", "pairs", "S2",
                    "pad rows + S3 + hand-over", "teardown", "t0: bulk issues", "zero fill", "expand scores", "row sums + image"]
        names_sr = ["issue", "wait free stage"]
        names_sm = ["issue / loop", "wait image", "wait accumulator", "wait ring stage", "MMA issue"]
        out["stack_fwd_consumer_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_sc, sf[:16])}
        out["stack_fwd_ring_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_sr, sf[16:32])}
        out["stack_fwd_mma_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_sm, sf[32:])}
    if hasattr(L, "molkgnn_debug_phase_clocks_bwdp"):
        bp = read("molkgnn_debug_phase_clocks_bwdp", 48)
        names_w = ["prologue", "wait tile data", "wait Wt buffer free", "Wt clear", "rank-0 scatter", "ranks 1-3 + hand-over",
                   "dxh epilogue (incl. waiting for the MMAs)", "G store"]
        names_r = ["issue", "wait free stage"]
        names_m = ["loop", "wait Wt", "wait ring stage", "MMA issue"]
        out["bwd_pipe_worker_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_w, bp[:16])}
        out["bwd_pipe_ring_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_r, bp[16:32])}
        out["bwd_pipe_mma_kcycles_per_cta_per_step"] = {n: v / ncta / steps / 1e3 for n, v in zip(names_m, bp[32:])}
# (2) plan reused: no host sync inside the loop -> wall time of the issue loop = host cost when the GPU is the bottleneck
plan = net.build_plan(t["edge_index"], t["p"], t["edge_attr"], t["x"].shape[0])
for _ in range(2):
    step(plan)
torch.cuda.synchronize()
w0 = time.perf_counter()
e0.record()
for _ in range(steps):
    step(plan)
e1.record()
w1 = time.perf_counter()
torch.cuda.synchronize()
w2 = time.perf_counter()
out["ms_per_step_plan_reused"] = e0.elapsed_time(e1) / steps
out["host_issue_ms_per_step_plan_reused"] = (w1 - w0) * 1e3 / steps
out["wall_ms_per_step_plan_reused"] = (w2 - w0) * 1e3 / steps
print(json.dumps(out, indent=1))
