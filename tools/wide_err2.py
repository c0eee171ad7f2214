"""Diagnostic: where does the wide backward's input gradient differ from the SIMT backward's (1 layer)?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth, _lib
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    L = (40, 80, 120, 200)
    b = synth.make_batch(nmol, seed=43)
    torch.manual_seed(43)
    net = mk.MolGCN(nl, *L, *L, x_dim=28, p_dim=3, edge_attr_dim=7).to("cuda")
    wout = torch.randn(b["x"].shape[0], 440, device="cuda")
    d = {k: torch.from_numpy(v).to("cuda") for k, v in b.items() if k in ("x", "p", "edge_index", "edge_attr")}
    res = {}
    for name, bp in (("simt", 0), ("wide", 1)):
        _lib.lib().molkgnn_set_bwd_path(bp)
        for p in net.parameters():
            p.grad = None
        x = d["x"].clone().requires_grad_(True)
        h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False)
        (h * wout).sum().backward()
        res[name] = (x.grad.double().cpu().numpy(), {n: p.grad.double().cpu().numpy() for n, p in net.named_parameters() if p.grad is not None})
    _lib.lib().molkgnn_set_bwd_path(1)
    a, r = res["wide"][0], res["simt"][0]
    deg = np.bincount(b["edge_index"][0], minlength=a.shape[0])
    err = np.abs(a - r)
    print("max|ref|", np.abs(r).max(), "max err", err.max(), "rel", err.max() / np.abs(r).max())
    print("mean err", err.mean(), "median", np.median(err))
    for dg in range(1, 5):
        m = deg == dg
        if m.any():
            print("deg", dg, "nodes", int(m.sum()), "max err", err[m].max(), "mean err", err[m].mean(), "mean |ref|", np.abs(r[m]).mean())
    print("per column max err", np.round(err.max(axis=0) / np.abs(r).max(), 8))
    idx = np.argsort(err.ravel())[::-1][:8]
    for i in idx:
        v, f = divmod(int(i), a.shape[1])
        print("node", v, "deg", int(deg[v]), "col", f, "wide", a[v, f], "simt", r[v, f], "rowmax", np.abs(r[v]).max())
    # signed bias: is the error systematic?
    sgn = np.sign(r) * (a - r)
    print("mean signed error (along ref sign)", sgn.mean(), "vs mean |err|", err.mean())
    for n in res["wide"][1]:
        ga, gr = res["wide"][1][n], res["simt"][1][n]
        e = np.abs(ga - gr).max() / max(np.abs(gr).max(), 1e-30)
        if e > 2e-6:
            print("param", n, "rel err", e)


if __name__ == "__main__":
    main()
