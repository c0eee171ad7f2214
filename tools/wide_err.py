"""Diagnostic: forward / input-gradient error of the wide model against the oracle, SIMT backward vs the tensor-core wide backward,
for 1..5 layers (teacher-forced arg-max).  Usage: python tools/wide_err.py [molecules]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth, _lib
    from tests.helpers import compact_from_kernel_major, rel_err
    from tests import test_conv_gpu as T
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    L = (40, 80, 120, 200)
    out = []
    for nl in (1, 2, 3, 5):
        b = synth.make_batch(nmol, seed=43)
        torch.manual_seed(43)
        net = mk.MolGCN(nl, *L, *L, x_dim=28, p_dim=3, edge_attr_dim=7)
        wout = torch.randn(b["x"].shape[0], 440)
        h_ref, gx_ref, params_ref, auxs = T._oracle_run(net, b, wout)
        net = net.to("cuda")
        d = T._to_dev(b)
        forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], "cuda") for aux in auxs]
        row = {"layers": nl}
        for name, fp, bp in (("simt_bwd", 3, 0), ("wide_bwd", 3, 1), ("simt_all", 0, 0)):
            _lib.lib().molkgnn_set_fwd_path(fp)
            _lib.lib().molkgnn_set_bwd_path(bp)
            for p in net.parameters():
                p.grad = None
            x = d["x"].clone().requires_grad_(True)
            h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, argmax_in=forced)
            (h * wout.to("cuda")).sum().backward()
            gerr = 0.0
            for li, layer in enumerate(net.layers):
                for dg, kc in enumerate(layer.trainable_kernelconv_set):
                    for nme in ("x_center", "x_support", "edge_attr_support"):
                        ref = params_ref[li][dg][nme].grad
                        if ref is not None and getattr(kc, nme).grad is not None:
                            gerr = max(gerr, rel_err(getattr(kc, nme).grad.cpu(), ref))
            row[name] = {"h": rel_err(h.detach().cpu(), h_ref), "gx": rel_err(x.grad.cpu(), gx_ref), "gparam_max": gerr}
        _lib.lib().molkgnn_set_fwd_path(3)
        _lib.lib().molkgnn_set_bwd_path(1)
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
