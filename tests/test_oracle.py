"""CPU: pin oracle/molkgnn_oracle.py against the fixtures produced by the UNMODIFIED reference (tools/make_golden.py)."""
import os
import numpy as np
import pytest
import torch

from oracle import molkgnn_oracle as orc
from tests.helpers import load_golden, golden_params, golden_buckets, golden_argmax, check_argmax, rel_err


def test_kat_cosine_docstring():
    # the reference's only in-tree known-answer vector, kernels.py:161-170 -> tensor([1.000, 0.8729])
    g = load_golden("kat_cosine")
    out = orc.cosine_mean(torch.from_numpy(g["t1"]), torch.from_numpy(g["t2"]))
    assert np.allclose(out.numpy(), g["out"], atol=1e-12)
    assert np.allclose(out.numpy(), [1.0, 0.8729], atol=5e-5)


def test_perm_tables():
    g = load_golden("perms")
    for d in range(1, 5):
        assert np.array_equal(np.asarray(orc.perm_table(d)), g[f"d{d}"])
    assert [len(orc.perm_table(d)) for d in range(1, 5)] == [1, 2, 6, 12]


@pytest.mark.parametrize("name", ["bucket_a", "bucket_b"])
def test_bucket_pass_bit_exact(name):
    g = load_golden(name)
    bk = orc.bucket_pass(g["edge_index"], int(g["num_nodes"]), g["p"], g["edge_attr"])
    for d in range(1, 5):
        for k in ["selected_index", "nei_index", "p_focal", "nei_p", "nei_edge_attr"]:
            ref = g[f"{k}_deg{d}"]
            got = bk[d][k]
            assert got.shape == ref.shape, (k, d, got.shape, ref.shape)
            assert np.array_equal(got, ref), (k, d)


@pytest.mark.parametrize("name", ["molgcn_small", "molgcn_readme", "molgcn_1layer", "molgcn_chains", "molgcn_stars"])
def test_molgcn_forward_backward(name):
    g = load_golden(name)
    params = golden_params(g, requires_grad=True)
    bk_np = orc.bucket_pass(g["edge_index"], g["x"].shape[0], g["p"], g["edge_attr"])
    bk = orc.buckets_to_torch(bk_np)
    gb = golden_buckets(g)
    for d in range(1, 5):  # oracle bucket pass == what the reference consumed
        for k in gb[d]:
            if gb[d][k].numel() == 0:          # a bucket that is empty in every molecule (wrapper.py:627-630): shapes are not pinned
                assert bk[d][k].numel() == 0
            else:
                assert torch.equal(bk[d][k], gb[d][k])
    x = torch.from_numpy(g["x"]).clone().requires_grad_(True)
    ei = torch.from_numpy(g["edge_index"])
    # teacher-forced on the reference's arg-max so that a flip inside a tie class cannot cascade (hard part 1)
    h, auxs = orc.molgcn_forward(params, x, ei, bk, return_aux=True, force_argmax=golden_argmax(g))
    assert rel_err(h.detach(), g["h"]) < 1e-5
    # arg-max permutation (tie-aware) against the reference's torch.max
    tot = ex = 0
    for li, aux in enumerate(auxs):
        for d in range(1, 5):
            if aux[d - 1] is None:
                continue
            n, e, _ = check_argmax(g[f"S_l{li}_d{d}"], g[f"argmax_l{li}_d{d}"], aux[d - 1]["argmax"])
            tot += n
            ex += e
    # bit-identical share outside the tie classes' free choice; the star fixture is ALL structural ties from layer 1 on (the four
    # leaves of a centre carry identical features, every permutation scores the same and rounding picks the winner)
    assert ex / tot > (0.75 if name == "molgcn_stars" else 0.97)
    (h * torch.from_numpy(g["wout"])).sum().backward()
    assert rel_err(x.grad, g["grad_x"]) < 2e-5
    for li, layer in enumerate(params):
        for d in range(4):
            if f"grad_layers.{li}.trainable_kernelconv_set.{d}.x_center" not in g:
                # empty bucket: the reference never ran this degree's KernelConv (kernels.py:690), autograd left .grad = None
                assert f"S_l{li}_d{d + 1}" not in g
                for nme in ["x_center", "x_support", "edge_attr_support"]:
                    assert layer[d][nme].grad is None or float(layer[d][nme].grad.abs().max()) == 0.0
                continue
            for nme in ["x_center", "x_support", "edge_attr_support"]:
                ref = g[f"grad_layers.{li}.trainable_kernelconv_set.{d}.{nme}"]
                assert rel_err(layer[d][nme].grad, ref) < 5e-5, (li, d, nme)
            trip = ["support_attr_sc_weight", "center_attr_sc_weight", "edge_attr_support_sc_weight"]
            refs = np.array([g[f"grad_layers.{li}.trainable_kernelconv_set.{d}.{t}"] for t in trip])
            got = np.array([layer[d][t].grad.item() for t in trip])
            assert np.abs(got - refs).max() <= 1e-4 * max(np.abs(refs).max(), 1e-6), (li, d, got, refs)
            # parameters the reference never differentiates
            assert f"grad_layers.{li}.trainable_kernelconv_set.{d}.p_support" not in g


def test_mixed_fixed_and_trainable_kernel_sets():
    """A layer with a fixed AND a trainable KernelConv per degree (kernels.py:702-715: rows of a degree = [fixed ; trainable]):
    the oracle, run once per set with the columns laid side by side, against the unmodified reference's BaseKernelSetConv."""
    from tests.helpers import PARAM_NAMES
    g = load_golden("set_mixed")

    def params(kind, rg):
        return [{n: torch.from_numpy(np.asarray(g[f"param_{kind}_kernelconv_set.{d}.{n}"])).clone().requires_grad_(rg)
                 for n in PARAM_NAMES} for d in range(4)]

    pf, pt = params("fixed", False), params("trainable", True)
    bk = orc.buckets_to_torch(orc.bucket_pass(g["edge_index"], g["x"].shape[0], g["p"], g["edge_attr"]))
    x = torch.from_numpy(g["x"]).clone().requires_grad_(True)
    sc_f = orc.kernel_set_conv_forward(pf, x, bk, is_last_layer=True)
    sc_t = orc.kernel_set_conv_forward(pt, x, bk, is_last_layer=True)
    cols, of, ot = [], 0, 0
    for d in range(4):
        Lf, Lt = int(g["Lf"][d]), int(g["Lt"][d])
        cols += [sc_f[:, of:of + Lf], sc_t[:, ot:ot + Lt]]
        of, ot = of + Lf, ot + Lt
    sc = torch.cat(cols, dim=1)
    assert rel_err(sc.detach(), g["sc"]) < 1e-5
    (sc * torch.from_numpy(g["wout"])).sum().backward()
    assert rel_err(x.grad, g["grad_x"]) < 2e-5
    for d in range(4):
        for nme in ["x_center", "x_support", "edge_attr_support"]:
            assert rel_err(pt[d][nme].grad, g[f"grad_trainable_kernelconv_set.{d}.{nme}"]) < 5e-5, (d, nme)
            assert f"grad_fixed_kernelconv_set.{d}.{nme}" not in g


@pytest.mark.skipif(not os.path.isdir("/root/reference/models/MolKGNN"), reason="the reference tree only exists in the build container")
def test_oracle_against_the_live_reference_on_fresh_batches():
    """Beyond the committed fixtures: tools/live_reference_check.py imports the UNMODIFIED reference (own process: it mocks
    rdkit & co. in sys.modules) and compares it with the oracle on fresh random batches -- random molecule counts, layer
    counts and kernel counts per seed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "live_reference_check.py"), "6"], capture_output=True,
                       text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["bucket_bit_exact"]
    assert out["h"] < 1e-5 and out["grad_x"] < 2e-5 and out["grad_params"] < 5e-5
    assert out["argmax_exact_share"] > 0.97


def test_oracle_at_the_molkgnnnet_call_site():
    """The reference's real call site (MolKGNNNet.py:115-146): BatchNorm1d on x and on edge_attr, conv stack on the RAW
    precomputed bond rows (kernels.py:679), swish-MLP, global_add_pool.  The fixture is the unmodified MolKGNNNet."""
    g = load_golden("molkgnnnet_call")
    sd = {k[len("param_"):]: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k.startswith("param_")}
    nl = int(g["num_layers"])
    params = []
    for i in range(nl):
        layer = []
        for d in range(4):
            layer.append({n: sd[f"gnn.layers.{i}.trainable_kernelconv_set.{d}.{n}"].clone().requires_grad_(True)
                          for n in ["x_center", "x_support", "edge_attr_support", "p_support", "length_sc_weight",
                                    "angle_sc_weight", "center_attr_sc_weight", "support_attr_sc_weight",
                                    "edge_attr_support_sc_weight"]})
        params.append(layer)
    x = torch.from_numpy(g["x"]).clone().requires_grad_(True)
    xb = torch.nn.functional.batch_norm(x, None, None, sd["node_batch_norm.weight"], sd["node_batch_norm.bias"], True)
    bk = golden_buckets(g)                       # raw bond rows / coordinates, as the reference consumed them
    h = orc.molgcn_forward(params, xb, torch.from_numpy(g["edge_index"]), bk)
    z = torch.nn.functional.linear(h, sd["graph_embedding_lin1.weight"], sd["graph_embedding_lin1.bias"])
    z = torch.nn.functional.linear(z * torch.sigmoid(z), sd["graph_embedding_lin2.weight"], sd["graph_embedding_lin2.bias"])
    batch = torch.from_numpy(g["batch"])
    out = torch.zeros(int(batch.max()) + 1, z.shape[1]).index_add(0, batch, z)
    (out * torch.from_numpy(g["wout"])).sum().backward()
    assert rel_err(out.detach(), g["out"]) < 1e-5
    assert rel_err(x.grad, g["grad_x"]) < 1e-5
    for i in range(nl):
        for d in range(4):
            for n in ("x_center", "x_support", "edge_attr_support"):
                assert rel_err(params[i][d][n].grad, g[f"grad_gnn.layers.{i}.trainable_kernelconv_set.{d}.{n}"]) < 1e-5
