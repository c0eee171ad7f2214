"""The drop-in at the reference's REAL call site (VERDICT r1 "what's weak" #1, ADVICE high).

``MolKGNNNet.forward`` (reference MolKGNNNet.py:115-119) batch-normalises ``edge_attr`` and hands THAT tensor to
``MolGCN.forward`` together with the 20 raw precomputed per-degree tensors; the reference conv reads only the latter
(kernels.py:679).  The fixture ``tests/golden/molkgnnnet_call.npz`` is the output / every gradient of the UNMODIFIED
reference MolKGNNNet (tools/make_golden.py, train mode, dropout 0).  Here the same network runs with ONLY ``MolGCN``
swapped for ``molkgnn_b200.MolGCN`` and NO extra keyword argument:

  * where ``oracle/_ref`` is staged (tools/make_oracle_ref.py; ships to the GPU box) the host network is the reference's
    own unmodified ``MolKGNNNet`` class with its module-level name ``MolGCN`` patched;
  * otherwise a line-by-line restatement of its forward (same attribute names, so the golden state dict loads).
"""
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

DEG_KEYS = ["p_focal", "nei_p", "nei_edge_attr", "selected_index", "nei_index"]


class Bag(object):
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def swish(x):
    return x * torch.sigmoid(x)


def add_pool(x, batch):
    n = int(batch.max()) + 1
    return torch.zeros(n, x.shape[1], dtype=x.dtype, device=x.device).index_add(0, batch, x)


class RestatedMolKGNNNet(torch.nn.Module):
    """MolKGNNNet.py:10-149 restated for boxes without oracle/_ref (attribute names = state-dict keys of the reference)."""

    def __init__(self, gnn, K, x_dim, edge_attr_dim, emb):
        super().__init__()
        self.graph_embedding_linear = torch.nn.Linear(K, emb)          # dead in the reference too (MolKGNNNet.py:20-25)
        self.node_batch_norm = torch.nn.BatchNorm1d(x_dim)
        self.edge_batch_norm = torch.nn.BatchNorm1d(edge_attr_dim)
        self.graph_embedding_lin1 = torch.nn.Linear(K, emb)
        self.graph_embedding_lin2 = torch.nn.Linear(emb, emb)
        self.gnn = gnn

    def forward(self, data, save_score=False):
        x = self.node_batch_norm(data.x)
        edge_attr = self.edge_batch_norm(data.edge_attr)               # MolKGNNNet.py:116: THIS goes to the gnn
        kw = {f"{k}_deg{d}": getattr(data, f"{k}_deg{d}") for d in range(1, 5) for k in DEG_KEYS}
        h = self.gnn(x=x, edge_index=data.edge_index, edge_attr=edge_attr, p=data.p, save_score=save_score, **kw)
        return add_pool(self.graph_embedding_lin2(swish(self.graph_embedding_lin1(h))), data.batch)


def build_net(g, dev):
    import molkgnn_b200 as mk
    import make_oracle_ref
    L1, LN = [int(v) for v in g["L1"]], [int(v) for v in g["LN"]]
    kw = dict(num_layers=int(g["num_layers"]), num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1], num_kernel3_1hop=L1[2],
              num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1], num_kernel3_Nhop=LN[2],
              num_kernel4_Nhop=LN[3], x_dim=g["x"].shape[1], p_dim=3, edge_attr_dim=g["edge_attr"].shape[1])
    if make_oracle_ref.available():
        mods = make_oracle_ref.load()
        host = mods["MolKGNNNet"]
        orig = host.MolGCN
        host.MolGCN = mk.MolGCN                      # the ONLY patch: the name MolKGNNNet.__init__ resolves (MolKGNNNet.py:1,45)
        try:
            net = host.MolKGNNNet(drop_ratio=0.0, graph_embedding_dim=int(g["emb"]), **kw)
        finally:
            host.MolGCN = orig
        kind = "reference MolKGNNNet (oracle/_ref)"
    else:
        net = RestatedMolKGNNNet(mk.MolGCN(**kw), sum(LN), g["x"].shape[1], g["edge_attr"].shape[1], int(g["emb"]))
        kind = "restated MolKGNNNet"
    assert isinstance(net.gnn, mk.MolGCN)
    sd = {k[len("param_"):]: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k.startswith("param_")}
    net.load_state_dict(sd, strict=True)
    return net.to(dev).train(), kind


def make_data(g, dev, requires_grad=True):
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(requires_grad)
    d = Bag(x=x, p=torch.from_numpy(g["p"]).to(dev), edge_index=torch.from_numpy(g["edge_index"]).to(dev),
            edge_attr=torch.from_numpy(g["edge_attr"]).to(dev), batch=torch.from_numpy(g["batch"]).to(dev))
    for dd in range(1, 5):
        for k in DEG_KEYS:
            setattr(d, f"{k}_deg{dd}", torch.from_numpy(g[f"bk_{k}_deg{dd}"]).to(dev))
    return d


def elementwise_ok(a, b, rtol=1e-5, floor_frac=1e-5):
    """|a - b| <= rtol * |b| + floor, floor = floor_frac * max|b| (the absolute floor of an fp32 sum whose terms are O(max|b|))"""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    floor = floor_frac * float(b.abs().max().clamp_min(1e-30))
    return bool(((a - b).abs() <= rtol * b.abs() + floor).all())


def test_molkgnnnet_callsite_matches_reference():
    g = load_golden("molkgnnnet_call")
    dev = torch.device("cuda", 0)
    net, kind = build_net(g, dev)
    data = make_data(g, dev)
    out = net(data)                                   # no raw_edge_attr=, no argmax_in=: the reference call, verbatim
    (out * torch.from_numpy(g["wout"]).to(dev)).sum().backward()
    torch.cuda.synchronize()
    errs = {"out": rel_err(out.detach().cpu(), g["out"]), "grad_x": rel_err(data.x.grad.cpu(), g["grad_x"])}
    assert elementwise_ok(out.detach().cpu(), g["out"]) and elementwise_ok(data.x.grad.cpu(), g["grad_x"])
    n_grads = 0
    for name, prm in net.named_parameters():
        key = "grad_" + name
        if key not in g:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, f"{name}: gradient where the reference has none"
            continue
        assert prm.grad is not None, f"{name}: no gradient"
        n_grads += 1
        if name.endswith("_sc_weight"):
            continue                                  # judged per (support, centre, edge) triple below
        errs[name] = rel_err(prm.grad.cpu(), g[key])
    # the three softmax mixing weights of a KernelConv: gradients sum to ~0, judged relative to the triple's max
    for li in range(int(g["num_layers"])):
        for d in range(4):
            base = f"gnn.layers.{li}.trainable_kernelconv_set.{d}."
            names = [base + n for n in ("support_attr_sc_weight", "center_attr_sc_weight", "edge_attr_support_sc_weight")]
            ref = np.array([float(g["grad_" + n]) for n in names])
            got = np.array([float(dict(net.named_parameters())[n].grad) for n in names])
            errs[base + "theta"] = float(np.abs(ref - got).max() / max(np.abs(ref).max(), 1e-30))
    worst = max(errs, key=errs.get)
    print(f"call site [{kind}]: {n_grads} gradients, min top-2 gap of the fixture {float(g['min_top2_gap']):.2e}, "
          f"worst rel err {errs[worst]:.2e} ({worst})")
    for k, v in errs.items():
        assert v < (1e-4 if k.endswith("theta") else 1e-5), f"{k}: {v:.3e}"


def test_callsite_fixture_is_sensitive_to_the_edge_attr_source():
    """Negative control: convolving the batch-normalised edge_attr (what round 1 did at this call site) is FAR from the
    reference, so the test above really pins the source of the bond rows."""
    g = load_golden("molkgnnnet_call")
    dev = torch.device("cuda", 0)
    net, _ = build_net(g, dev)
    data = make_data(g, dev, requires_grad=False)
    with torch.no_grad():
        x = net.node_batch_norm(data.x)
        ea_bn = net.edge_batch_norm(data.edge_attr)
        kw = {f"{k}_deg{d}": getattr(data, f"{k}_deg{d}") for d in range(1, 5) for k in DEG_KEYS}
        h_ref_protocol = net.gnn(x=x, edge_index=data.edge_index, edge_attr=ea_bn, p=data.p, save_score=False, **kw)
        h_raw = net.gnn(x=x, edge_index=data.edge_index, edge_attr=data.edge_attr, p=data.p, save_score=False)
        h_bn = net.gnn(x=x, edge_index=data.edge_index, edge_attr=ea_bn, p=data.p, save_score=False)
    assert rel_err(h_ref_protocol, h_raw) < 1e-6          # raw tensors from kwargs == gather of the raw edge_attr
    assert rel_err(h_bn, h_raw) > 1e-3                    # BatchNorm'd rows change the edge term


def test_inconsistent_reference_tensors_raise():
    import molkgnn_b200 as mk
    g = load_golden("molkgnnnet_call")
    dev = torch.device("cuda", 0)
    net, _ = build_net(g, dev)
    data = make_data(g, dev, requires_grad=False)
    kw = {f"{k}_deg{d}": getattr(data, f"{k}_deg{d}") for d in range(1, 5) for k in DEG_KEYS}
    kw["nei_edge_attr_deg2"] = kw["nei_edge_attr_deg2"][:-1]
    with pytest.raises(mk.MolKGNNError), torch.no_grad():
        net.gnn(x=data.x, edge_index=data.edge_index, edge_attr=data.edge_attr, p=data.p, save_score=False, **kw)


def test_save_score_passthrough(tmp_path, monkeypatch):
    """MolGCN.forward(save_score=True) reaches every layer's save_score (KernelLayer.py:117, kernels.py:749-750, 594-608):
    scores.csv holds the LAST layer's sim_sc [K, N] (each layer rewrites the file), equal to the oracle's."""
    import pandas as pd
    from oracle import molkgnn_oracle as orc
    from tests.helpers import module_from_golden, params_from_module, golden_buckets
    g = load_golden("molgcn_small")
    dev = torch.device("cuda", 0)
    net = module_from_golden(g, dev)
    monkeypatch.chdir(tmp_path)
    os.makedirs("customized_kernels")
    for d in range(1, 5):                                 # one (empty) list of predefined kernel names per degree
        pd.DataFrame({"name": []}).to_csv(f"customized_kernels/deg{d}.csv", index=False)
    kw = dict(x=torch.from_numpy(g["x"]).to(dev), edge_index=torch.from_numpy(g["edge_index"]).to(dev),
              edge_attr=torch.from_numpy(g["edge_attr"]).to(dev), p=torch.from_numpy(g["p"]).to(dev))
    with torch.no_grad():
        h = net(save_score=True, **kw)
        h2 = net(save_score=False, **kw)
    assert torch.equal(h, h2)
    got = pd.read_csv("scores.csv", index_col=0).to_numpy()
    K = net.layers[-1].get_num_kernel()
    assert got.shape == (K, g["x"].shape[0])
    # oracle: sim_sc of the last layer
    params = params_from_module(net)
    bk = golden_buckets(g)
    x = torch.from_numpy(g["x"])
    ei = torch.from_numpy(g["edge_index"])
    hh = x
    for li, lp in enumerate(params):
        sc = orc.kernel_set_conv_forward(lp, hh, bk, is_last_layer=(li == len(params) - 1))
        hh = orc.propagate(ei, sc)
    # free-running arg-max on a fixture WITH structural ties (duplicated leaves): a flipped tie changes single entries by O(0.1)
    # (the edge term of another permutation), everything else is exact to rounding
    close = np.abs(got.T - sc.numpy()) <= 1e-5 * np.abs(sc.numpy()).max()
    assert close.mean() > 0.97, close.mean()
    assert np.median(np.abs(got.T - sc.numpy())) < 1e-6


def test_backward_after_repack_with_modified_parameters_is_refused():
    """forward A, in-place parameter update, forward B, backward A: the packed kernel rows are one workspace per module
    (ADVICE r1): the stale backward must raise instead of differentiating with B's kernels."""
    import molkgnn_b200 as mk
    from tests.helpers import module_from_golden
    g = load_golden("molgcn_small")
    dev = torch.device("cuda", 0)
    net = module_from_golden(g, dev)
    kw = dict(edge_index=torch.from_numpy(g["edge_index"]).to(dev), edge_attr=torch.from_numpy(g["edge_attr"]).to(dev),
              p=torch.from_numpy(g["p"]).to(dev), save_score=False)
    xa = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    ha = net(x=xa, **kw)
    hb_same = net(x=xa, **kw)                      # a second forward with UNCHANGED parameters is fine
    ha.sum().backward()
    ha = net(x=xa, **kw)
    with torch.no_grad():
        net.layers[0].trainable_kernelconv_set[1].x_center.mul_(1.5)
    net(x=xa, **kw)                                # re-packs with the modified parameters
    with pytest.raises(mk.MolKGNNError):
        ha.sum().backward()
    del hb_same


def test_molgcn_with_mixed_fixed_and_trainable_layer_and_save_kernels(tmp_path):
    """a MolGCN whose layers carry fixed AND trainable kernel sets (kernels.py:452-516, 702-715) runs through
    the drop-in (layer by layer) and equals the oracle; save_kernels writes the '{deg-1}.{param}' keys kernel_reader.py:86 reads."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from oracle import molkgnn_oracle as orc
    from tests.helpers import PARAM_NAMES
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    b = synth.make_batch(12, seed=33, dup_leaf_prob=0.0)
    net = mk.MolGCN(2, 3, 4, 5, 6, 2, 3, 4, 5, x_dim=28, p_dim=3, edge_attr_dim=7)
    Lf = (2, 1, 2, 1)
    fixed = [mk.KernelConv(L=Lf[d], D=3, num_supports=d + 1, node_attr_dim=28, edge_attr_dim=7, requires_grad=False,
                           weight_requires_grad=False) for d in range(4)]
    train = list(net.layers[0].trainable_kernelconv_set)
    net.layers[0] = mk.BaseKernelSetConv(*fixed, *train)
    K0 = sum(Lf) + 3 + 4 + 5 + 6
    net.layers[1] = mk.KernelSetConv(2, 3, 4, 5, D=3, node_attr_dim=K0, edge_attr_dim=7)
    net = net.to(dev)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("x", "p", "edge_index", "edge_attr")}
    x = t["x"].clone().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    h.sum().backward()
    # oracle: a degree's score rows are [fixed ; trainable] (kernels.py:702-715)
    bk = orc.buckets_to_torch(orc.bucket_pass(b["edge_index"], b["x"].shape[0], b["p"], b["edge_attr"]))
    cpu = lambda kc: {n: getattr(kc, n).detach().cpu() for n in PARAM_NAMES}  # noqa: E731
    xr = torch.from_numpy(b["x"])
    ei = torch.from_numpy(b["edge_index"])
    sc_f = orc.kernel_set_conv_forward([cpu(k) for k in fixed], xr, bk, is_last_layer=False)
    sc_t = orc.kernel_set_conv_forward([cpu(k) for k in train], xr, bk, is_last_layer=False)
    cols, of, ot = [], 0, 0
    Lt = (3, 4, 5, 6)
    for d in range(4):
        cols += [sc_f[:, of:of + Lf[d]], sc_t[:, ot:ot + Lt[d]]]
        of, ot = of + Lf[d], ot + Lt[d]
    h1 = orc.propagate(ei, torch.cat(cols, dim=1))
    sc2 = orc.kernel_set_conv_forward([cpu(k) for k in net.layers[1].trainable_kernelconv_set], h1, bk, is_last_layer=True)
    h_ref = orc.propagate(ei, sc2)
    close = (h.detach().cpu() - h_ref).abs() <= 1e-5 * h_ref.abs().max()
    assert close.float().mean() > 0.97          # free-running arg-max: all but tie flips
    assert x.grad is not None and all(p.grad is None for k in fixed for p in k.parameters())
    net.save_kernels(str(tmp_path) + "/", "kernels.pt")
    sd = torch.load(str(tmp_path) + "/kernels.pt")
    assert set(sd) == {f"{d}.{n}" for d in range(4) for n in PARAM_NAMES}
    assert sd["3.x_support"].shape == (6, 4, 28)


def test_native_molkgnnnet_matches_the_reference_fixture():
    """molkgnn_b200.MolKGNNNet (same ctor / state dict / forward protocol as MolKGNNNet.py:10-149, native conv
    stack, deterministic native global_add_pool) against the fixture of the unmodified reference network."""
    import molkgnn_b200 as mk
    g = load_golden("molkgnnnet_call")
    dev = torch.device("cuda", 0)
    L1, LN = [int(v) for v in g["L1"]], [int(v) for v in g["LN"]]
    net = mk.MolKGNNNet(num_layers=int(g["num_layers"]), num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1], num_kernel3_1hop=L1[2],
                        num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1], num_kernel3_Nhop=LN[2],
                        num_kernel4_Nhop=LN[3], x_dim=g["x"].shape[1], p_dim=3, edge_attr_dim=g["edge_attr"].shape[1],
                        drop_ratio=0.0, graph_embedding_dim=int(g["emb"]))
    sd = {k[len("param_"):]: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k.startswith("param_")}
    net.load_state_dict(sd, strict=True)               # identical state-dict keys
    net = net.to(dev).train()
    data = make_data(g, dev)
    out = net(data)
    (out * torch.from_numpy(g["wout"]).to(dev)).sum().backward()
    assert rel_err(out.detach().cpu(), g["out"]) < 1e-5
    assert rel_err(data.x.grad.cpu(), g["grad_x"]) < 1e-5
    for name, prm in net.named_parameters():
        if "grad_" + name in g and not name.endswith("_sc_weight"):
            assert rel_err(prm.grad.cpu(), g["grad_" + name]) < 1e-5, name
    # the same network on a batch WITHOUT the precomputed per-degree tensors (molkgnn_b200.store batches): same output
    bare = Bag(x=data.x.detach(), p=data.p, edge_index=data.edge_index, edge_attr=data.edge_attr, batch=data.batch)
    net.zero_grad()
    with torch.no_grad():                              # (torch.no_grad does not switch BatchNorm to eval: still batch statistics)
        out2 = net(bare)
    assert rel_err(out2.cpu(), g["out"]) < 1e-5
    # running statistics of BOTH batch norms advanced, like the reference's (the edge one is otherwise dead)
    old = torch.from_numpy(g["param_edge_batch_norm.running_mean"])
    mean = torch.from_numpy(g["edge_attr"]).mean(0)
    assert rel_err(net.edge_batch_norm.running_mean.cpu(), 0.9 * (0.9 * old + 0.1 * mean) + 0.1 * mean) < 1e-5   # two train-mode calls
    with pytest.raises(ValueError):
        net(data, data)
