"""Shared helpers for the parity tests: golden loading, parameter plumbing, tie-aware argmax comparison."""
import os

import numpy as np
import torch

from oracle import molkgnn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAM_NAMES = ["x_center", "x_support", "edge_attr_support", "p_support", "length_sc_weight", "angle_sc_weight",
               "center_attr_sc_weight", "support_attr_sc_weight", "edge_attr_support_sc_weight"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_params(g, dtype=torch.float32, requires_grad=False, prefix="param_"):
    """state-dict keys 'layers.{i}.trainable_kernelconv_set.{d-1}.{name}' -> [[{name: tensor} x4] x layers]."""
    nl = int(g["num_layers"])
    out = []
    for i in range(nl):
        layer = []
        for d in range(4):
            prm = {}
            for n in PARAM_NAMES:
                key = f"{prefix}layers.{i}.trainable_kernelconv_set.{d}.{n}"
                if key in g:
                    prm[n] = torch.from_numpy(np.asarray(g[key])).to(dtype).clone().requires_grad_(requires_grad)
            layer.append(prm)
        out.append(layer)
    return out


def golden_buckets(g):
    bk = {}
    for d in range(1, 5):
        bk[d] = {k: torch.from_numpy(np.asarray(g[f"bk_{k}_deg{d}"])) for k in
                 ["selected_index", "nei_index", "p_focal", "nei_p", "nei_edge_attr"]}
        bk[d]["selected_index"] = bk[d]["selected_index"].long()
        bk[d]["nei_index"] = bk[d]["nei_index"].long()
    return bk


def check_argmax(S_ref, am_ref, am_test, gap_tol=1e-5):
    """Tie-aware arg-max check (hard part 1).

    S_ref [L,P,n]: permutation scores as the reference computed them; am_ref/am_test [L,n].
    Requires identical index wherever the reference's top-2 gap >= gap_tol; otherwise the tested index must lie in
    the reference's tie class (score within gap_tol of the max).  Returns (n_total, n_exact, n_in_tie_class)."""
    S = torch.as_tensor(S_ref).double()
    am_ref = torch.as_tensor(am_ref).long()
    am_test = torch.as_tensor(am_test).long()
    best = S.max(dim=1).values
    chosen = torch.gather(S, 1, am_test.unsqueeze(1)).squeeze(1)
    exact = (am_ref == am_test)
    in_class = (best - chosen) <= gap_tol
    bad = ~(exact | in_class)
    assert not bad.any(), f"{int(bad.sum())} arg-max entries outside the reference's tie class"
    return am_ref.numel(), int(exact.sum()), int((in_class & ~exact).sum())


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def golden_argmax(g):
    """[[argmax [L_d,n_d] or None x4] x layers] as the reference's torch.max chose (kernels.py:373)."""
    out = []
    for li in range(int(g["num_layers"])):
        out.append([torch.from_numpy(g[f"argmax_l{li}_d{d}"]) if f"argmax_l{li}_d{d}" in g else None
                    for d in range(1, 5)])
    return out
