#!/usr/bin/env python
"""Live parity of the CPU oracle against the UNMODIFIED reference on fresh random batches (beyond the committed fixtures).

Runs only where /root/reference exists (the build container); imports the reference the way tools/make_golden.py does.
For every seed: the reference's per-molecule transform + collation vs the oracle's bucket pass (bit-exact), and the reference's
MolGCN forward / autograd backward vs the oracle teacher-forced on the reference's arg-max.  Prints one JSON line.

    python tools/live_reference_check.py [n_seeds]
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)                      # puts the stub + /root/reference on sys.path, imports the reference

from oracle import molkgnn_oracle as orc  # noqa: E402
from molkgnn_b200 import synth  # noqa: E402
from tests.helpers import rel_err, check_argmax, PARAM_NAMES  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out = {"seeds": n_seeds, "bucket_bit_exact": True, "h": 0.0, "grad_x": 0.0, "grad_params": 0.0, "argmax_exact_share": 1.0,
       "cases": []}
for s in range(n_seeds):
    seed = 100 + s
    rng = np.random.default_rng(seed)
    n_mol = int(rng.integers(2, 6))
    layers = int(rng.integers(1, 4))
    L1 = tuple(int(v) for v in rng.integers(1, 7, 4))
    LN = tuple(int(v) for v in rng.integers(1, 7, 4))
    mols = synth.make_molecules(n_mol, seed=seed, dup_leaf_prob=float(rng.choice([0.0, 0.5])))
    b = synth.collate(mols)
    bk_ref = mg.ref_bucket_collated(mols)
    N = b["x"].shape[0]
    bk_np = orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"])
    for d in range(1, 5):
        for k in mg.DEG_KEYS:
            ref, got = bk_ref[f"{k}_deg{d}"], bk_np[d][k]
            if ref.size == 0:
                out["bucket_bit_exact"] &= got.size == 0
            else:
                out["bucket_bit_exact"] &= bool(got.shape == ref.shape and np.array_equal(got, ref))
    torch.manual_seed(seed)
    net = mg.MolGCN(num_layers=layers, num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1], num_kernel3_1hop=L1[2],
                    num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1], num_kernel3_Nhop=LN[2],
                    num_kernel4_Nhop=LN[3], x_dim=synth.X_DIM, p_dim=3, edge_attr_dim=synth.EDGE_DIM)
    x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    kw = dict(x=x, edge_index=torch.from_numpy(b["edge_index"]), edge_attr=torch.from_numpy(b["edge_attr"]),
              p=torch.from_numpy(b["p"]), save_score=False)
    for k, v in bk_ref.items():
        kw[k] = torch.from_numpy(v)
    with mg.MaxSpy() as spy:
        h = net(**kw)
    wout = torch.randn(h.shape, generator=torch.Generator().manual_seed(seed))
    (h * wout).sum().backward()
    # oracle, teacher-forced on the reference's arg-max
    it = iter(spy.records)
    forced, S_ref = [], []
    for li in range(layers):
        fl, sl = [], []
        for d in range(1, 5):
            if bk_ref[f"selected_index_deg{d}"].size == 0:
                fl.append(None); sl.append(None)
                continue
            S, am = next(it)
            fl.append(am); sl.append(S)
        forced.append(fl); S_ref.append(sl)
    params = [[{n: getattr(kc, n).detach().clone().requires_grad_(True) for n in PARAM_NAMES}
               for kc in layer.trainable_kernelconv_set] for layer in net.layers]
    xo = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    ho, auxs = orc.molgcn_forward(params, xo, torch.from_numpy(b["edge_index"]), orc.buckets_to_torch(bk_np),
                                  return_aux=True, force_argmax=forced)
    (ho * wout).sum().backward()
    tot = ex = 0
    for li in range(layers):
        for d in range(4):
            if auxs[li][d] is None:
                continue
            n_, e_, _ = check_argmax(S_ref[li][d], forced[li][d], auxs[li][d]["argmax"])   # raises outside the tie class
            tot += n_; ex += e_
    e_h, e_x = rel_err(ho.detach(), h.detach()), rel_err(xo.grad, x.grad)
    e_p = 0.0
    for li, layer in enumerate(net.layers):
        for d, kc in enumerate(layer.trainable_kernelconv_set):
            for n in ("x_center", "x_support", "edge_attr_support"):
                if getattr(kc, n).grad is not None:
                    e_p = max(e_p, rel_err(params[li][d][n].grad, getattr(kc, n).grad))
            for n in ("p_support", "length_sc_weight", "angle_sc_weight"):
                assert getattr(kc, n).grad is None
    out["h"], out["grad_x"], out["grad_params"] = max(out["h"], e_h), max(out["grad_x"], e_x), max(out["grad_params"], e_p)
    out["argmax_exact_share"] = min(out["argmax_exact_share"], ex / max(tot, 1))
    out["cases"].append(dict(seed=seed, molecules=n_mol, layers=layers, L1=L1, LN=LN, N=N))
print(json.dumps(out))
