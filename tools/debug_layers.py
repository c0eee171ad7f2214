"""Temporary debugging aid: per-layer comparison of CUDA intermediates against the oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import molkgnn_oracle as orc
from tests.helpers import *
from molkgnn_b200 import functional as Fn
from molkgnn_b200.plan import BucketPlan

name = sys.argv[1] if len(sys.argv) > 1 else "molgcn_readme"
g = load_golden(name)
dev = "cuda"
net = module_from_golden(g, dev)
params = golden_params(g, requires_grad=True)
N = g["x"].shape[0]
bk = orc.buckets_to_torch(orc.bucket_pass(g["edge_index"], N, g["p"], g["edge_attr"]))
ei = torch.from_numpy(g["edge_index"])
# oracle with retained per-layer inputs
x0 = torch.from_numpy(g["x"]).clone().requires_grad_(True)
hs = [x0]
h = x0
scs = []
am = golden_argmax(g)
for i, lp in enumerate(params):
    sc = orc.kernel_set_conv_forward(lp, h, bk, is_last_layer=(i == len(params) - 1), force_argmax=am[i])
    sc.retain_grad(); scs.append(sc)
    h = orc.propagate(ei, sc); h.retain_grad(); hs.append(h)
(h * torch.from_numpy(g["wout"])).sum().backward()

plan = BucketPlan.from_edge_index(ei.to(dev), torch.from_numpy(g["p"]).to(dev), torch.from_numpy(g["edge_attr"]).to(dev), N)
layer_params = [layer._degree_params() for layer in net.layers]
packs, saved = [], []
F = 28
hh, hn = None, None
for i, prm in enumerate(layer_params):
    pack = Fn.LayerPack(prm, F, 7, dev).pack()
    if i == 0:
        hh, hn = Fn.pad_norm(torch.from_numpy(g["x"]).to(dev), pack.Fp)
    forced = compact_from_kernel_major(am[i], dev)
    sc, argmax, free = Fn.conv_forward(plan, pack, hh, hn, i == len(layer_params) - 1, dense=False, argmax_in=forced)
    saved.append((hh, hn, argmax))
    hh, hn = Fn.propagate_forward(plan, pack, sc)
    print("fwd layer", i, "h err", rel_err(hh[:, :pack.K].cpu(), hs[i + 1].detach()))
    packs.append(pack); F = pack.K
gr = torch.from_numpy(g["wout"]).to(dev)
for i in range(len(packs) - 1, -1, -1):
    xp, xn, argmax = saved[i]
    gx, grads = Fn.conv_backward(plan, packs[i], xp, xn, gr, 1, argmax, True, True)
    ref = hs[i].grad
    err = (gx[:, :packs[i].F].cpu() - ref).abs()
    print("bwd layer", i, "gx err", float(err.max() / ref.abs().max()), "worst rows", err.max(1).values.topk(5).indices.tolist(),
          "deg of worst", [int(bk['deg'][r]) for r in err.max(1).values.topk(5).indices.tolist()])
    for d in range(4):
        for nme in ["x_center", "x_support", "edge_attr_support"]:
            e = rel_err(grads[d][nme].cpu(), params[i][d][nme].grad)
            if e > 1e-5: print("   param", d + 1, nme, e)
    gr = gx

# ---- focus: last layer gx ----
i = len(packs) - 1
xp, xn, argmax = saved[i]
gx, grads = Fn.conv_backward(plan, packs[i], xp, xn, torch.from_numpy(g["wout"]).to(dev), 1, argmax, True, False)
ref = hs[i].grad
err = (gx[:, :packs[i].F].cpu() - ref).abs() / ref.abs().max()
print("col err max per 8-col block:", [f"{float(err[:, c:c+8].max()):.1e}" for c in range(0, 110, 8)])
bad = (err.max(1).values > 1e-5).nonzero().flatten().tolist()
print("bad rows", len(bad), "of", N, bad[:40])
print("deg of bad rows", np.bincount([int(bk['deg'][r]) for r in bad], minlength=5))
nbdeg = {}
src, dst = g["edge_index"]
for r in bad[:10]:
    print(r, "deg", int(bk['deg'][r]), "nbr degs", [int(bk['deg'][u]) for u in dst[src == r]], "err", float(err[r].max()))
good = [r for r in range(N) if r not in bad][:10]
for r in good:
    print("good", r, "deg", int(bk['deg'][r]), "nbr degs", [int(bk['deg'][u]) for u in dst[src == r]])

print("---- per degree-block isolation, last layer ----")
offs = np.concatenate([[0], np.cumsum(packs[i].L)])
for d in range(4):
    wmask = torch.zeros_like(torch.from_numpy(g["wout"]))
    wmask[:, offs[d]:offs[d + 1]] = torch.from_numpy(g["wout"])[:, offs[d]:offs[d + 1]]
    hin = hs[i].detach().clone().requires_grad_(True)
    sc = orc.kernel_set_conv_forward(params[i], hin, bk, is_last_layer=True, force_argmax=am[i])
    hout = orc.propagate(ei, sc)
    (hout * wmask).sum().backward()
    gx, _ = Fn.conv_backward(plan, packs[i], xp, xn, wmask.to(dev), 1, argmax, True, False)
    err = (gx[:, :packs[i].F].cpu() - hin.grad).abs() / hin.grad.abs().max()
    bad = (err.max(1).values > 1e-5).nonzero().flatten().tolist()
    print("block deg", d + 1, "max err", float(err.max()), "bad rows", bad[:10])
# node 1 neighbourhood details
src, dst = g["edge_index"]
print("node1 out-nbrs", dst[src == 1], "in_src", plan.in_src[1].tolist(), "in_j", plan.in_j[1].tolist(), "in_cnt", int(plan.in_cnt[1]))
for u in dst[src == 1]:
    print("  nbr", u, "deg", int(bk['deg'][u]), "its out-nbrs", dst[src == u].tolist(), "pos", int(plan.pos[u]))
xr = hs[i].detach()
for u in dst[src == 1]:
    nb = dst[src == u]
    for a_ in range(len(nb)):
        for b_ in range(a_ + 1, len(nb)):
            if torch.equal(xr[nb[a_]], xr[nb[b_]]): print("  dup rows among nbrs of", u, ":", nb[a_], nb[b_])
