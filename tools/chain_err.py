"""Diagnostic: kernel-parameter / input gradients of the tensor-core backward against the fp32 SIMT backward as the batch grows
(the tensor core truncates every accumulate: long accumulation chains in one TMEM accumulator drift).
Usage: python tools/chain_err.py [molecules ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import molkgnn_b200 as mk
    from molkgnn_b200 import _lib
    from configs_bench import make_batch
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]
    for B in sizes:
        b = make_batch(B)
        torch.manual_seed(0)
        net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to("cuda")
        d = {k: torch.from_numpy(b[k]).to("cuda") for k in ("x", "p", "edge_index", "edge_attr")}
        wout = torch.randn(d["x"].shape[0], 110, device="cuda")
        res = {}
        for name, bp in (("simt", 0), ("tile", 1)):
            _lib.lib().molkgnn_set_bwd_path(bp)
            _lib.lib().molkgnn_set_fwd_path(2)      # the same (per-layer) forward for both: identical arg-max choices
            for p in net.parameters():
                p.grad = None
            x = d["x"].clone().requires_grad_(True)
            h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False)
            (h * wout).sum().backward()
            res[name] = (x.grad.double().cpu().numpy(), {n: p.grad.double().cpu().numpy() for n, p in net.named_parameters() if p.grad is not None})
        _lib.lib().molkgnn_set_bwd_path(1)
        _lib.lib().molkgnn_set_fwd_path(3)
        gx = np.abs(res["tile"][0] - res["simt"][0]).max() / np.abs(res["simt"][0]).max()
        worst, wn, bias = 0.0, "", 0.0
        for n in res["tile"][1]:
            a, r = res["tile"][1][n], res["simt"][1][n]
            if a.size < 8:
                continue
            e = np.abs(a - r).max() / max(np.abs(r).max(), 1e-30)
            if e > worst:
                worst, wn = e, n
                bias = float((np.sign(r) * (a - r)).mean() / np.abs(r).mean())
        print(json.dumps({"molecules": B, "gx_rel": gx, "gparam_rel_max": worst, "param": wn, "mean_signed_rel": bias}), flush=True)


if __name__ == "__main__":
    main()
