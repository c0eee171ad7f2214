#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

Runs only where /root/reference exists (the build container).  The reference is imported through the
test-only ``tests/stubs/torch_geometric`` stand-in (torch_geometric is not installed, no network);
``rdkit`` & co. are mocked because only ``wrapper.ToXAndPAndEdgeAttrForDeg`` (pure torch) is executed.

    python tools/make_golden.py            # rewrites tests/golden/

Fixtures
  kat_cosine.npz        the reference's only in-tree known-answer vector (kernels.py:161-170), evaluated by
                        KernelConv.calculate_average_similarity_score itself
  perms.npz             KernelConv.permute applied to an index tensor -> the permutation tables (kernels.py:109-128)
  bucket_*.npz          ToXAndPAndEdgeAttrForDeg per molecule + PyG collation semantics (wrapper.py:559-672)
  molgcn_*.npz          MolGCN forward (+ torch.max spy: S tensor / argmax per layer & degree) and autograd grads
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
for m in ["rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.RDLogger", "rdkit.Chem.EState",
          "rdkit.Chem.rdMolDescriptors", "rdkit.Chem.rdPartialCharges", "rdkit.Chem.Crippen", "networkx",
          "models.ChIRoNet.embedding_functions", "clearml"]:
    sys.modules.setdefault(m, MagicMock())

from torch_geometric.data import Data  # noqa: E402  (the stub)
from models.MolKGNN.kernels import KernelConv  # noqa: E402
from models.MolKGNN.KernelLayer import MolGCN  # noqa: E402
import wrapper as ref_wrapper  # noqa: E402

from molkgnn_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ONLY = None
DEG_KEYS = ["p_focal", "nei_p", "nei_edge_attr", "selected_index", "nei_index"]


def ref_bucket_collated(mols):
    """Reference transform per molecule, then PyG 2.0.x collation semantics for the 20 per-degree attributes."""
    tr = ref_wrapper.ToXAndPAndEdgeAttrForDeg()
    per = []
    for m in mols:
        d = Data(x=torch.from_numpy(m.x), p=torch.from_numpy(m.p), edge_index=torch.from_numpy(m.edge_index),
                 edge_attr=torch.from_numpy(m.edge_attr))
        per.append(tr(d))
    out = {}
    off = np.concatenate([[0], np.cumsum([m.num_nodes for m in mols])])
    for d in range(1, 5):
        for key in DEG_KEYS:
            name = f"{key}_deg{d}"
            parts = []
            for i, dd in enumerate(per):
                t = getattr(dd, name)
                if t.numel() == 0:
                    continue  # PyG cat of an empty 1-D float tensor is a no-op
                if "index" in key:
                    t = t + int(off[i])
                parts.append(t)
            if parts:
                cat = torch.cat(parts, dim=-1 if "index" in key else 0)
            else:
                # every molecule's bucket is empty: the indices stay int64 (nonzero() / .to(torch.long), wrapper.py:600,632),
                # the float attributes are the empty float tensor of wrapper.py:627-630
                cat = torch.zeros(0, dtype=torch.int64) if "index" in key else torch.zeros(0)
            out[name] = cat.numpy()
    return out


class MaxSpy(object):
    def __init__(self):
        self.records = []
        self._orig = torch.max

    def __enter__(self):
        def spy(*a, **k):
            r = self._orig(*a, **k)
            if len(a) == 1 and k.get("dim", None) == 1 and a[0].dim() == 3:
                self.records.append((a[0].detach().clone(), r[1].detach().clone()))
            return r
        torch.max = spy
        return self

    def __exit__(self, *a):
        torch.max = self._orig


def mols_from_bonds(specs, seed):
    """hand-built molecules: list of (n_atoms, [(i, j), ...]); bond b -> edge rows 2b, 2b+1 (wrapper.py:152-156)"""
    rng = np.random.default_rng(seed)
    out = []
    for n, bonds in specs:
        x = rng.standard_normal((n, synth.X_DIM)).astype(np.float32)
        p = (rng.standard_normal((n, 3)) * 1.5).astype(np.float32)
        ei, ea = [], []
        for (i, j) in bonds:
            a = np.zeros(synth.EDGE_DIM, np.float32)
            a[rng.integers(0, 4)] = 1.0
            a[4:] = rng.integers(0, 2, 3)
            ei += [(i, j), (j, i)]
            ea += [a, a]
        out.append(synth.Molecule(x, p, np.array(ei, dtype=np.int64).T.copy(), np.stack(ea)))
    return out


def molgcn_case(name, n_mol, seed, num_layers, L1, LN, dup=0.5, mols=None):
    if mols is None:
        mols = synth.make_molecules(n_mol, seed=seed, dup_leaf_prob=dup)
    b = synth.collate(mols)
    bk = ref_bucket_collated(mols)
    torch.manual_seed(seed)
    net = MolGCN(num_layers=num_layers, num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1], num_kernel3_1hop=L1[2],
                 num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1], num_kernel3_Nhop=LN[2],
                 num_kernel4_Nhop=LN[3], x_dim=synth.X_DIM, p_dim=3, edge_attr_dim=synth.EDGE_DIM)
    # de-symmetrise the scalar mixing weights so their gradients are exercised away from the 0.2/0.2/0.2 init
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for layer in net.layers:
            for kc in layer.trainable_kernelconv_set:
                for w in (kc.support_attr_sc_weight, kc.center_attr_sc_weight, kc.edge_attr_support_sc_weight):
                    w.add_(0.5 * torch.randn((), generator=g))
    x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    kw = dict(x=x, edge_index=torch.from_numpy(b["edge_index"]), edge_attr=torch.from_numpy(b["edge_attr"]),
              p=torch.from_numpy(b["p"]), save_score=False)
    for k, v in bk.items():
        kw[k] = torch.from_numpy(v)
    with MaxSpy() as spy:
        h = net(**kw)
    K = h.shape[1]
    wout = torch.randn(h.shape, generator=torch.Generator().manual_seed(seed + 11))
    (h * wout).sum().backward()
    save = dict(x=b["x"], p=b["p"], edge_index=b["edge_index"], edge_attr=b["edge_attr"], batch=b["batch"],
                h=h.detach().numpy(), wout=wout.numpy(), grad_x=x.grad.numpy(),
                num_layers=np.int64(num_layers), L1=np.asarray(L1), LN=np.asarray(LN), seed=np.int64(seed))
    for k, v in bk.items():
        save["bk_" + k] = v
    for k, v in net.state_dict().items():
        save["param_" + k] = v.numpy()
    for k, v in net.named_parameters():
        if v.grad is not None:
            save["grad_" + k] = v.grad.numpy()
    # spy records arrive in (layer, degree) order for the non-empty buckets
    it = iter(spy.records)
    for li in range(num_layers):
        for d in range(1, 5):
            if bk[f"selected_index_deg{d}"].size == 0:
                continue
            S, am = next(it)
            save[f"S_l{li}_d{d}"] = S.numpy()          # [L,P,n_d] fp32, as the reference computed it
            save[f"argmax_l{li}_d{d}"] = am.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "N=", b["x"].shape[0], "E=", b["edge_index"].shape[1], "K=", K, "h.abs.mean=", float(h.abs().mean()))


def mixed_set_case(name, n_mol, seed, Lf, Lt):
    """One BaseKernelSetConv layer with a FIXED (requires_grad=False) and a TRAINABLE KernelConv per degree
    (kernels.py:702-715: the score rows of a degree are [fixed ; trainable]), last layer (chirality on)."""
    from models.MolKGNN.kernels import BaseKernelSetConv
    # no duplicated leaf rows: without structural ties the free-running arg-max is the reference's (a mixed layer runs as two
    # passes and cannot be teacher-forced)
    mols = synth.make_molecules(n_mol, seed=seed, dup_leaf_prob=0.0)
    b = synth.collate(mols)
    bk = ref_bucket_collated(mols)
    torch.manual_seed(seed)
    mk_kc = lambda L, d, rg: KernelConv(L=L, D=3, num_supports=d, node_attr_dim=synth.X_DIM,  # noqa: E731
                                        edge_attr_dim=synth.EDGE_DIM, requires_grad=rg, weight_requires_grad=rg)
    fixed = [mk_kc(Lf[d], d + 1, False) for d in range(4)]
    train = [mk_kc(Lt[d], d + 1, True) for d in range(4)]
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for kc in fixed + train:
            for w in (kc.support_attr_sc_weight, kc.center_attr_sc_weight, kc.edge_attr_support_sc_weight):
                w.add_(0.5 * torch.randn((), generator=g))
    layer = BaseKernelSetConv(*fixed, *train)
    x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    data = Data(x=x, edge_index=torch.from_numpy(b["edge_index"]), edge_attr=torch.from_numpy(b["edge_attr"]),
                p=torch.from_numpy(b["p"]), **{k: torch.from_numpy(v) for k, v in bk.items()})
    sc = layer(is_last_layer=True, data=data, save_score=False)
    wout = torch.randn(sc.shape, generator=torch.Generator().manual_seed(seed + 11))
    (sc * wout).sum().backward()
    save = dict(x=b["x"], p=b["p"], edge_index=b["edge_index"], edge_attr=b["edge_attr"], sc=sc.detach().numpy(),
                wout=wout.numpy(), grad_x=x.grad.numpy(), Lf=np.asarray(Lf), Lt=np.asarray(Lt), seed=np.int64(seed))
    for k, v in bk.items():
        save["bk_" + k] = v
    for k, v in layer.state_dict().items():
        save["param_" + k] = v.numpy()
    for k, v in layer.named_parameters():
        if v.grad is not None:
            save["grad_" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "N=", b["x"].shape[0], "K=", sc.shape[1], "grads:", sum(1 for k in save if k.startswith("grad_")))


def no_sibling_leaves(m):
    """True if no two degree-1 atoms of the molecule share their neighbour (such leaves carry bit-identical features from
    layer 1 on, which ties permutations structurally -- SURVEY.md 7 hard part 1)."""
    src, dst = m.edge_index
    deg = np.bincount(src, minlength=m.num_nodes)
    parents = [int(dst[np.nonzero(src == v)[0][0]]) for v in range(m.num_nodes) if deg[v] == 1]
    return len(parents) == len(set(parents))


def molkgnnnet_case(name, n_mol, seed, num_layers, L1, LN, emb=32):
    """The reference's REAL call site of the hot path: the unmodified MolKGNNNet (MolKGNNNet.py:69-149) -- BatchNorm1d on x
    AND on edge_attr (the latter is what reaches MolGCN.forward as `edge_attr`, MolKGNNNet.py:115-119), the conv stack, the
    swish-MLP and global_add_pool.  Train mode (batch statistics), dropout 0 so the run is deterministic.  Molecules without
    sibling leaves, so that the free-running arg-max has no structural ties: the fixture records the smallest top-2 gap."""
    from models.MolKGNN.MolKGNNNet import MolKGNNNet
    cand = synth.make_molecules(8 * n_mol, seed=seed, dup_leaf_prob=0.0)
    mols = [m for m in cand if no_sibling_leaves(m)][:n_mol]
    assert len(mols) == n_mol, "not enough molecules without sibling leaves"
    b = synth.collate(mols)
    bk = ref_bucket_collated(mols)
    torch.manual_seed(seed)
    net = MolKGNNNet(num_layers=num_layers, num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1], num_kernel3_1hop=L1[2],
                     num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1], num_kernel3_Nhop=LN[2],
                     num_kernel4_Nhop=LN[3], x_dim=synth.X_DIM, p_dim=3, edge_attr_dim=synth.EDGE_DIM, drop_ratio=0.0,
                     graph_embedding_dim=emb)
    net.train()
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for layer in net.gnn.layers:
            for kc in layer.trainable_kernelconv_set:
                for w in (kc.support_attr_sc_weight, kc.center_attr_sc_weight, kc.edge_attr_support_sc_weight):
                    w.add_(0.5 * torch.randn((), generator=g))
        net.node_batch_norm.weight.add_(0.3 * torch.randn(synth.X_DIM, generator=g))
        net.node_batch_norm.bias.add_(0.3 * torch.randn(synth.X_DIM, generator=g))
        net.edge_batch_norm.weight.add_(0.3 * torch.randn(synth.EDGE_DIM, generator=g))
        net.edge_batch_norm.bias.add_(0.3 * torch.randn(synth.EDGE_DIM, generator=g))
    x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    data = Data(x=x, p=torch.from_numpy(b["p"]), edge_index=torch.from_numpy(b["edge_index"]),
                edge_attr=torch.from_numpy(b["edge_attr"]), batch=torch.from_numpy(b["batch"]),
                **{k: torch.from_numpy(v) for k, v in bk.items()})
    with MaxSpy() as spy:
        out = net(data)
    wout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 11))
    (out * wout).sum().backward()
    gap = np.inf
    for S, am in spy.records:
        if S.shape[1] > 1:
            top2 = torch.topk(S.double(), 2, dim=1).values
            gap = min(gap, float((top2[:, 0] - top2[:, 1]).min()))
    save = dict(x=b["x"], p=b["p"], edge_index=b["edge_index"], edge_attr=b["edge_attr"], batch=b["batch"],
                out=out.detach().numpy(), wout=wout.numpy(), grad_x=x.grad.numpy(), num_layers=np.int64(num_layers),
                L1=np.asarray(L1), LN=np.asarray(LN), seed=np.int64(seed), emb=np.int64(emb), min_top2_gap=np.float64(gap))
    for k, v in bk.items():
        save["bk_" + k] = v
    for k, v in net.state_dict().items():
        save["param_" + k] = v.numpy()
    for k, v in net.named_parameters():
        if v.grad is not None:
            save["grad_" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "N=", b["x"].shape[0], "graphs=", out.shape[0], "min top-2 gap=", gap,
          "grads:", sum(1 for k in save if k.startswith("grad_")))


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. docstring KAT (kernels.py:161-170)
    kc = KernelConv(L=1, D=3, num_supports=1, node_attr_dim=3, edge_attr_dim=1)
    t1 = torch.tensor([[[1, 2, 3], [3, 2, 1]], [[1, 2, 3], [3, 2, 1]]], dtype=torch.double)
    t2 = torch.tensor([[[1, 2, 3], [3, 2, 1]], [[1, 2, 1], [1, 2, 1]]], dtype=torch.double)
    r = kc.calculate_average_similarity_score(t1, t2, sim_dim=-1, avg_dim=-2)
    np.savez(os.path.join(OUT, "kat_cosine.npz"), t1=t1.numpy(), t2=t2.numpy(), out=r.numpy())
    print("KAT", r)
    # 2. permutation tables
    tabs = {}
    for d in range(1, 5):
        idx = torch.arange(d).view(1, d, 1)
        tabs[f"d{d}"] = kc.permute(idx)[0, :, :, 0].numpy()
    np.savez(os.path.join(OUT, "perms.npz"), **tabs)
    # 3. bucket transform
    for name, n_mol, seed in [("bucket_a", 12, 3), ("bucket_b", 5, 17)]:
        mols = synth.make_molecules(n_mol, seed=seed)
        b = synth.collate(mols)
        bk = ref_bucket_collated(mols)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), p=b["p"], edge_index=b["edge_index"],
                            edge_attr=b["edge_attr"], num_nodes=np.int64(b["x"].shape[0]), **bk)
        print(name, {k: v.shape for k, v in bk.items() if "selected" in k})
    # 4. MolGCN forward/backward
    molgcn_case("molgcn_small", n_mol=6, seed=1, num_layers=2, L1=(3, 4, 5, 6), LN=(2, 3, 4, 5))
    molgcn_case("molgcn_readme", n_mol=4, seed=2, num_layers=3, L1=(10, 20, 30, 50), LN=(10, 20, 30, 50))
    molgcn_case("molgcn_1layer", n_mol=5, seed=5, num_layers=1, L1=(4, 4, 4, 4), LN=(4, 4, 4, 4))
    # 5. empty degree buckets (kernels.py:702-721 skips them; wrapper.py:627-630 stores empty tensors)
    chains = mols_from_bonds([(n, [(i, i + 1) for i in range(n - 1)]) for n in (2, 5, 9, 2, 3)], 73)
    molgcn_case("molgcn_chains", None, 73, num_layers=2, L1=(3, 4, 5, 6), LN=(2, 3, 4, 5), mols=chains)
    stars = mols_from_bonds([(5, [(0, 1), (0, 2), (0, 3), (0, 4)]), (2, [(0, 1)]),
                             (8, [(0, 1), (0, 2), (0, 3), (0, 4), (4, 5), (4, 6), (4, 7)])], 74)
    molgcn_case("molgcn_stars", None, 74, num_layers=3, L1=(10, 20, 30, 50), LN=(10, 20, 30, 50), mols=stars)
    # 6. fixed + trainable kernel sets in one layer (kernels.py:452-516, 702-715)
    mixed_set_case("set_mixed", n_mol=5, seed=9, Lf=(2, 3, 2, 3), Lt=(3, 2, 4, 2))
    # 7. the real call site: unmodified MolKGNNNet around the stack (BatchNorm'd edge_attr reaches MolGCN.forward)
    if ONLY in (None, "molkgnnnet"):
        molkgnnnet_case("molkgnnnet_call", n_mol=6, seed=29, num_layers=3, L1=(10, 20, 30, 50), LN=(10, 20, 30, 50))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "molkgnnnet":      # only the new fixture (the others are unchanged)
        ONLY = "molkgnnnet"
        os.makedirs(OUT, exist_ok=True)
        molkgnnnet_case("molkgnnnet_call", n_mol=6, seed=29, num_layers=3, L1=(10, 20, 30, 50), LN=(10, 20, 30, 50))
    else:
        main()
