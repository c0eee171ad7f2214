"""GPU parity tests proper: the CUDA conv stack (through the C-ABI, via the drop-in modules) against
  (a) the golden fixtures produced by the UNMODIFIED reference (tests/golden, tools/make_golden.py) and
  (b) the CPU oracle on fresh seeded inputs.

Tolerances (BASELINE.json north_star): degree buckets and arg-max permutation indices exact -- arg-max up to the
reference's own tie classes (hard part 1: structurally tied permutations are decided by fp32 rounding of
ATen's dot-product order, which no other implementation can reproduce); scores and gradients within 1e-5 relative (fp32),
measured as max |err| / max |ref| per tensor.  Mixing-weight gradients (which cancel to ~0 across the softmax triple)
are judged relative to the largest magnitude of their triple at 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import molkgnn_oracle as orc
from tests.helpers import (load_golden, golden_argmax, check_argmax, rel_err, compact_from_kernel_major,
                           kernel_major_from_compact, module_from_golden, params_from_module)

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


PATHS = {"simt": 0, "bucket_tc": 1, "tile": 2, "fused": 3}


@pytest.fixture(autouse=True, params=["fused", "tile", "bucket_tc", "simt"])
def fwd_path(request):
    """Every test runs four times: with the layer-fused molecule-tile tcgen05 forward (default product path: the whole stack in
    one launch), the per-layer molecule-tile forward, the bucket-order tcgen05 forward (plans without tiles) and the fp32 SIMT
    forward (wide layers)."""
    from molkgnn_b200 import _lib
    old = _lib.lib().molkgnn_set_fwd_path(PATHS[request.param])
    oldb = _lib.lib().molkgnn_set_bwd_path(1 if request.param in ("tile", "fused") else 0)   # tile backward with the tile forwards
    yield request.param
    _lib.lib().molkgnn_set_fwd_path(3 if old < 0 else old)
    _lib.lib().molkgnn_set_bwd_path(oldb)


def _to_dev(b):
    return dict(x=torch.from_numpy(b["x"]).to(DEV), p=torch.from_numpy(b["p"]).to(DEV),
                edge_index=torch.from_numpy(b["edge_index"]).to(DEV), edge_attr=torch.from_numpy(b["edge_attr"]).to(DEV))


def _check_param_grads(net, ref_grad, tol=TOL):
    """ref_grad(li, d, name) -> reference gradient (numpy / tensor), or None where the reference has none: a degree bucket
    that is empty in the whole batch never runs its KernelConv (kernels.py:690), so autograd leaves .grad = None there and the
    CUDA path must deliver None or exact zeros"""
    for li, layer in enumerate(net.layers):
        for d, kc in enumerate(layer.trainable_kernelconv_set):
            trip = ["support_attr_sc_weight", "center_attr_sc_weight", "edge_attr_support_sc_weight"]
            if ref_grad(li, d, "x_center") is None:
                for nme in ["x_center", "x_support", "edge_attr_support"] + trip:
                    got = getattr(kc, nme).grad
                    assert got is None or float(got.abs().max()) == 0.0, (li, d, nme)
                continue
            for nme in ["x_center", "x_support", "edge_attr_support"]:
                got = getattr(kc, nme).grad
                assert got is not None, (li, d, nme)
                assert rel_err(got.cpu(), ref_grad(li, d, nme)) < tol, (li, d, nme)
            refs = np.array([float(ref_grad(li, d, t)) for t in trip])
            got = np.array([getattr(kc, t).grad.item() for t in trip])
            assert np.abs(got - refs).max() <= 1e-4 * max(np.abs(refs).max(), 1e-6), (li, d, got, refs)
            # never differentiated by the reference
            for nme in ["p_support", "length_sc_weight", "angle_sc_weight"]:
                assert getattr(kc, nme).grad is None


def test_tile_kernels_are_the_default_path(fwd_path):
    """The base model on a tiled plan must run the molecule-tile tcgen05 kernels (no silent fall back)."""
    if fwd_path not in ("tile", "fused"):
        pytest.skip("path selection is only asserted for the tile paths")
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth, functional
    b = _to_dev(synth.make_batch(64, seed=11))
    torch.manual_seed(0)
    net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(DEV)
    before = functional.path_counts()
    x = b["x"].clone().requires_grad_(True)
    h = net(x=x, edge_index=b["edge_index"], edge_attr=b["edge_attr"], p=b["p"], save_score=False)
    h.sum().backward()
    torch.cuda.synchronize()
    after = functional.path_counts()
    assert after["fwd_tile"] - before["fwd_tile"] == 3 and after["fwd_other"] == before["fwd_other"]
    assert after["bwd_tile"] - before["bwd_tile"] == 3 and after["bwd_other"] == before["bwd_other"]




@pytest.mark.parametrize("name", ["molgcn_small", "molgcn_readme", "molgcn_1layer", "molgcn_chains", "molgcn_stars"])
def test_molgcn_vs_reference_golden(name):
    g = load_golden(name)
    net = module_from_golden(g, DEV)
    b = _to_dev(g)
    x = b["x"].clone().requires_grad_(True)
    plan = net.build_plan(b["edge_index"], b["p"], b["edge_attr"], x.shape[0])
    forced = [compact_from_kernel_major(a, DEV) for a in golden_argmax(g)]
    aux = {}
    h = net(x=x, edge_index=b["edge_index"], edge_attr=b["edge_attr"], p=b["p"], save_score=False, plan=plan,
            argmax_in=forced, aux=aux)
    assert h.shape == g["h"].shape
    assert rel_err(h.detach().cpu(), g["h"]) < TOL
    # free-running arg-max of every layer/degree vs the reference's torch.max (layer inputs are teacher-forced)
    tot = ex = 0
    for li, layer in enumerate(net.layers):
        free = kernel_major_from_compact(aux["argmax_free"][li], plan.n, layer.num_kernel_list)
        used = kernel_major_from_compact(aux["argmax"][li], plan.n, layer.num_kernel_list)
        for d in range(1, 5):
            if free[d - 1] is None:
                continue
            n_, e_, _ = check_argmax(g[f"S_l{li}_d{d}"], g[f"argmax_l{li}_d{d}"], free[d - 1])
            tot += n_
            ex += e_
            assert torch.equal(used[d - 1] & 0x7f, torch.from_numpy(g[f"argmax_l{li}_d{d}"]).to(torch.uint8))
    assert ex / tot > (0.75 if name == "molgcn_stars" else 0.97), (ex, tot)   # stars: all structural ties
    (h * torch.from_numpy(g["wout"]).to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), g["grad_x"]) < TOL
    _check_param_grads(net, lambda li, d, n: g[f"grad_layers.{li}.trainable_kernelconv_set.{d}.{n}"]
                       if f"grad_layers.{li}.trainable_kernelconv_set.{d}.{n}" in g else None)


def _oracle_run(net, b, wout, force=None):
    params = params_from_module(net, requires_grad=True)
    N = b["x"].shape[0]
    bk = orc.buckets_to_torch(orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"]))
    x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
    h, auxs = orc.molgcn_forward(params, x, torch.from_numpy(b["edge_index"]), bk, return_aux=True, force_argmax=force)
    (h * wout).sum().backward()
    return h.detach(), x.grad, params, auxs


@pytest.mark.parametrize("n_mol,layers,L1,LN,seed", [
    (48, 3, (10, 20, 30, 50), (10, 20, 30, 50), 11),      # README shape, a few tiles per degree
    (300, 2, (3, 5, 7, 9), (4, 6, 8, 10), 12),            # odd kernel counts: partial kernel groups, K % 4 != 0
    (16, 2, (1, 1, 1, 1), (1, 2, 1, 2), 13),              # minimum kernel counts
    (1, 3, (10, 20, 30, 50), (10, 20, 30, 50), 21),       # a single molecule: one tile, one CTA walks every kernel block
    (3, 2, (10, 20, 30, 50), (10, 20, 30, 50), 22),       # fewer tiles than kernel blocks
])
def test_molgcn_vs_oracle(n_mol, layers, L1, LN, seed):
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    b = synth.make_batch(n_mol, seed=seed)
    torch.manual_seed(seed)
    net = mk.MolGCN(layers, *L1, *LN, x_dim=28, p_dim=3, edge_attr_dim=7)
    with torch.no_grad():
        for layer in net.layers:
            for kc in layer.trainable_kernelconv_set:
                for w in (kc.support_attr_sc_weight, kc.center_attr_sc_weight, kc.edge_attr_support_sc_weight):
                    w.add_(0.5 * torch.randn(()))
    wout = torch.randn(b["x"].shape[0], sum(LN) if layers > 1 else sum(L1))
    # oracle free-running first: its arg-max is then forced onto the GPU run (tie classes checked separately)
    h_ref, gx_ref, params_ref, auxs = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    x = d["x"].clone().requires_grad_(True)
    plan = net.build_plan(d["edge_index"], d["p"], d["edge_attr"], x.shape[0])
    forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
    aux = {}
    h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, plan=plan,
            argmax_in=forced, aux=aux)
    assert rel_err(h.detach().cpu(), h_ref) < TOL
    tot = ex = 0
    for li, layer in enumerate(net.layers):
        free = kernel_major_from_compact(aux["argmax_free"][li], plan.n, layer.num_kernel_list)
        for dd in range(1, 5):
            if free[dd - 1] is None:
                continue
            n_, e_, _ = check_argmax(auxs[li][dd - 1]["S"].detach(), auxs[li][dd - 1]["argmax"], free[dd - 1])
            tot += n_
            ex += e_
    assert ex / tot > 0.97, (ex, tot)
    (h * wout.to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), gx_ref) < TOL
    _check_param_grads(net, lambda li, dg, n: params_ref[li][dg][n].grad)


def test_free_running_is_deterministic_and_close():
    """Without teacher forcing the run is deterministic (bitwise equal twice) and differs from the oracle only
    downstream of tie flips: the bulk of the output matches to 1e-5."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    b = synth.make_batch(64, seed=21)
    torch.manual_seed(21)
    net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7)
    wout = torch.randn(b["x"].shape[0], 110)
    h_ref, gx_ref, _, _ = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    outs = []
    for _ in range(2):
        x = d["x"].clone().requires_grad_(True)
        h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False)
        (h * wout.to(DEV)).sum().backward()
        outs.append((h.detach().clone(), x.grad.clone(), [p.grad.clone() for p in net.parameters() if p.grad is not None]))
        net.zero_grad()
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    for a, c in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, c)
    # tie flips cascade through later layers, so only coarse closeness is asserted here; exact parity is covered by
    # the teacher-forced tests above
    err = (outs[0][0].cpu() - h_ref).abs() / h_ref.abs().max()
    assert float(err.mean()) < 1e-2


def test_kernel_set_conv_layer_dense():
    """KernelSetConv.forward with the reference protocol (data= bag carrying the precomputed per-degree tensors)."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from molkgnn_b200.kernels import _Bag
    b = synth.make_batch(40, seed=31)
    N = b["x"].shape[0]
    bk_np = orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"])
    bk = orc.buckets_to_torch(bk_np)
    torch.manual_seed(3)
    layer = mk.KernelSetConv(4, 6, 8, 10, D=3, node_attr_dim=28, edge_attr_dim=7)
    lp = [{n: getattr(kc, n).detach().clone().requires_grad_(True) for n in
           ["x_center", "x_support", "edge_attr_support", "p_support", "support_attr_sc_weight",
            "center_attr_sc_weight", "edge_attr_support_sc_weight"]} for kc in layer.trainable_kernelconv_set]
    for is_last in (False, True):
        xr = torch.from_numpy(b["x"]).clone().requires_grad_(True)
        sc_ref, aux_ref = orc.kernel_set_conv_forward(lp, xr, bk, is_last_layer=is_last, return_aux=True)
        layer = layer.to(DEV)
        data = _Bag(x=torch.from_numpy(b["x"]).to(DEV).requires_grad_(True), p=torch.from_numpy(b["p"]).to(DEV),
                    edge_index=torch.from_numpy(b["edge_index"]).to(DEV),
                    edge_attr=torch.from_numpy(b["edge_attr"]).to(DEV))
        for d in range(1, 5):
            for k, v in bk[d].items():
                setattr(data, f"{k}_deg{d}", v.to(DEV))
        forced = compact_from_kernel_major([None if a is None else a["argmax"] for a in aux_ref], DEV)
        sc = layer(is_last_layer=is_last, data=data, save_score=False, argmax_in=forced)
        assert sc.shape == (N, 28)
        assert rel_err(sc.detach().cpu(), sc_ref.detach()) < TOL
        # block sparsity: a degree-d node is non-zero only inside its own degree block (kernels.py:725-727)
        deg = torch.from_numpy(bk_np["deg"])
        offs = [0, 4, 10, 18, 28]
        for d in range(1, 5):
            rows = sc.detach().cpu()[deg == d]
            mask = torch.ones(28, dtype=torch.bool)
            mask[offs[d - 1]:offs[d]] = False
            assert (rows[:, mask] == 0).all()
        w = torch.randn(N, 28, generator=torch.Generator().manual_seed(1))
        (sc * w.to(DEV)).sum().backward()
        (sc_ref * w).sum().backward()
        assert rel_err(data.x.grad.cpu(), xr.grad) < TOL
        for d, kc in enumerate(layer.trainable_kernelconv_set):
            for nme in ["x_center", "x_support", "edge_attr_support"]:
                assert rel_err(getattr(kc, nme).grad.cpu(), lp[d][nme].grad) < TOL, (d, nme)
        layer.zero_grad()
        for prm in lp:
            for v in prm.values():
                v.grad = None


@pytest.mark.parametrize("deg", [1, 2, 3, 4])
def test_kernel_conv_single_bucket(deg):
    """KernelConv.forward, the reference's per-degree entry (kernels.py:428-448): kwargs protocol, [L,n] output."""
    import molkgnn_b200 as mk
    torch.manual_seed(deg)
    n, L, F = 37, 7, 28
    kc = mk.KernelConv(L=L, D=3, num_supports=deg, node_attr_dim=F, edge_attr_dim=7)
    prm = {k: getattr(kc, k).detach().clone() for k in
           ["x_center", "x_support", "edge_attr_support", "p_support", "support_attr_sc_weight",
            "center_attr_sc_weight", "edge_attr_support_sc_weight"]}
    x_focal, p_focal = torch.randn(n, F), torch.randn(n, 3)
    x_nei, p_nei = torch.randn(n, deg, F), torch.randn(n, deg, 3)
    if deg >= 2:
        x_nei[::5, 1] = x_nei[::5, 0]   # duplicate neighbours: tied permutations + chirality gate
    ea = torch.rand(n, deg, 7) + 0.1
    for is_last in (False, True):
        ref, aux = orc.kernel_conv_forward(prm, x_focal, p_focal, x_nei, p_nei, ea, is_last_layer=is_last,
                                           return_aux=True)
        kc = kc.to(DEV)
        out = kc(is_last_layer=is_last, x_focal=x_focal.to(DEV), p_focal=p_focal.to(DEV), x_neighbor=x_nei.to(DEV),
                 p_neighbor=p_nei.to(DEV), edge_attr_neighbor=ea.to(DEV))
        assert out.shape == (L, n)
        # free-running here: compare only where the oracle's arg-max gap is clear
        S = aux["S"].double()
        top2 = S.topk(min(2, S.shape[1]), dim=1).values
        clear = (top2[:, 0] - top2[:, -1] > 1e-5) if S.shape[1] > 1 else torch.ones_like(top2[:, 0], dtype=torch.bool)
        err = (out.detach().cpu() - ref).abs() / ref.abs().max()
        assert (err[clear] < TOL).all()
    with pytest.raises(Exception):
        kc(is_last_layer=False, x_focal=x_focal.to(DEV), p_focal=torch.zeros(n, 2, device=DEV),
           x_neighbor=x_nei.to(DEV), p_neighbor=p_nei.to(DEV), edge_attr_neighbor=ea.to(DEV))


def test_wide_layer_chunked_paths():
    """Wide kernel sets (BASELINE config 3 shape, fewer layers/molecules): kernel set does not fit in shared memory,
    so the feature-chunked forward, kernel-ranged backward and multi-pass input-gradient paths are exercised."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    b = synth.make_batch(6, seed=41)
    torch.manual_seed(41)
    L = (40, 80, 120, 200)
    net = mk.MolGCN(2, *L, *L, x_dim=28, p_dim=3, edge_attr_dim=7)
    wout = torch.randn(b["x"].shape[0], 440)
    h_ref, gx_ref, params_ref, auxs = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    x = d["x"].clone().requires_grad_(True)
    forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
    h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, argmax_in=forced)
    assert rel_err(h.detach().cpu(), h_ref) < TOL
    (h * wout.to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), gx_ref) < TOL
    _check_param_grads(net, lambda li, dg, n: params_ref[li][dg][n].grad)


def test_wide_five_layer_stack():
    """BASELINE configs[2] at a size the oracle finishes in seconds: kernels 40/80/120/200, FIVE layers (F = 28, then 440),
    so the last-layer chirality, the propagate between wide layers and the layer loop of the wide stack are all covered."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    b = synth.make_batch(24, seed=43)
    torch.manual_seed(43)
    L = (40, 80, 120, 200)
    net = mk.MolGCN(5, *L, *L, x_dim=28, p_dim=3, edge_attr_dim=7)
    wout = torch.randn(b["x"].shape[0], 440)
    h_ref, gx_ref, params_ref, auxs = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    x = d["x"].clone().requires_grad_(True)
    forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
    h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, argmax_in=forced)
    assert rel_err(h.detach().cpu(), h_ref) < TOL
    (h * wout.to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), gx_ref) < TOL
    _check_param_grads(net, lambda li, dg, n: params_ref[li][dg][n].grad)


def test_inference_sweep_forward_only():
    """BASELINE configs[4] (virtual-screening sweep): the forward under torch.no_grad() is the same kernels as the training
    forward -- bitwise the same h -- chunk after chunk through one module, and matches the oracle."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    torch.manual_seed(5)
    net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7)
    chunks = [synth.make_batch(n, seed=50 + n) for n in (40, 7, 64)]
    refs = []
    for b in chunks:
        h_ref, _, _, auxs = _oracle_run(net, b, torch.zeros(b["x"].shape[0], 110))
        refs.append((h_ref, auxs))
    net = net.to(DEV)
    for b, (h_ref, auxs) in zip(chunks, refs):
        d = _to_dev(b)
        with torch.no_grad():
            h0 = net(x=d["x"], edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False)
        assert not h0.requires_grad
        x = d["x"].clone().requires_grad_(True)
        h1 = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False)
        assert torch.equal(h0, h1.detach())
        forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
        with torch.no_grad():
            hf = net(x=d["x"], edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False,
                     argmax_in=forced)
        assert rel_err(hf.cpu(), h_ref) < TOL


def _batch_from_bonds(mols, seed):
    """collated batch from a list of (n_atoms, [(i, j), ...]) molecules; bond b occupies edge rows 2b, 2b+1 (wrapper.py:152-156)"""
    rng = np.random.default_rng(seed)
    xs, ps, ei, ea, off = [], [], [], [], 0
    for n, bonds in mols:
        xs.append(rng.standard_normal((n, 28)).astype(np.float32))
        ps.append((rng.standard_normal((n, 3)) * 1.5).astype(np.float32))
        for (i, j) in bonds:
            a = np.zeros(7, np.float32)
            a[rng.integers(0, 4)] = 1.0
            a[4:] = rng.integers(0, 2, 3)
            ei += [(off + i, off + j), (off + j, off + i)]
            ea += [a, a]
        off += n
    return dict(x=np.concatenate(xs), p=np.concatenate(ps), edge_index=np.array(ei, dtype=np.int64).T.copy(),
                edge_attr=np.stack(ea))


def _edge_case(name):
    from molkgnn_b200 import synth
    if name == "large_molecules":          # molecules that do not fit a 128-node tile: no tiling, bucket-order kernels
        return synth.make_batch(3, seed=71, min_atoms=140, max_atoms=150)
    if name == "molecules_of_128":             # one molecule fills a tile exactly
        return synth.make_batch(2, seed=72, min_atoms=128, max_atoms=128)
    if name == "mixed_sizes":                  # 20 .. 128 atoms: tiles of very different fill
        return synth.make_batch(12, seed=76, min_atoms=20, max_atoms=128)
    if name == "chains":                   # degrees 1 and 2 only: the degree-3 and degree-4 buckets are empty
        return _batch_from_bonds([(n, [(i, i + 1) for i in range(n - 1)]) for n in (2, 5, 9, 2, 3)], 73)
    if name == "stars":                    # degrees 1 and 4 only (+ a two-atom molecule)
        return _batch_from_bonds([(5, [(0, 1), (0, 2), (0, 3), (0, 4)]), (2, [(0, 1)]),
                                  (8, [(0, 1), (0, 2), (0, 3), (0, 4), (4, 5), (4, 6), (4, 7)])], 74)
    if name == "two_atoms":                # the smallest legal graph
        return _batch_from_bonds([(2, [(0, 1)])], 75)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["large_molecules", "molecules_of_128", "mixed_sizes", "chains", "stars", "two_atoms"])
def test_edge_case_graphs(name, fwd_path):
    """Graph shapes at the edges of the bucket / tile logic, each against the oracle (forced arg-max, 1e-5): molecules larger
    than a tile (the plan carries no tiles and the bucket-order kernels run), molecules that fill a tile exactly, mixed
    sizes up to 128 atoms, batches with empty degree buckets (kernels.py:702-721 skips them), the two-atom molecule."""
    import molkgnn_b200 as mk
    b = _edge_case(name)
    torch.manual_seed(7)
    net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7)
    wout = torch.randn(b["x"].shape[0], 110)
    h_ref, gx_ref, params_ref, auxs = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    x = d["x"].clone().requires_grad_(True)
    plan = net.build_plan(d["edge_index"], d["p"], d["edge_attr"], x.shape[0])
    if name == "large_molecules":
        assert plan.n_tiles == 0
    else:
        assert plan.n_tiles > 0
    if name == "molecules_of_128":
        assert plan.c.tile_max_nodes == 128 and plan.n_tiles == 2
    forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
    h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, plan=plan,
            argmax_in=forced)
    assert rel_err(h.detach().cpu(), h_ref) < TOL
    (h * wout.to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), gx_ref) < TOL
    _check_param_grads(net, lambda li, dg, n: params_ref[li][dg][n].grad)


def test_mixed_fixed_and_trainable_kernel_sets():
    """BaseKernelSetConv with a fixed (requires_grad=False) AND a trainable KernelConv per degree (kernels.py:452-516,
    702-715) against the unmodified reference (tests/golden/set_mixed.npz): score columns of a degree are [fixed ; trainable],
    gradients reach x and the trainable kernels only."""
    import types
    import molkgnn_b200 as mk
    g = load_golden("set_mixed")
    Lf, Lt = [int(v) for v in g["Lf"]], [int(v) for v in g["Lt"]]
    mk_kc = lambda L, d, rg: mk.KernelConv(L=L, D=3, num_supports=d, node_attr_dim=28, edge_attr_dim=7,  # noqa: E731
                                           requires_grad=rg, weight_requires_grad=rg)
    layer = mk.BaseKernelSetConv(*[mk_kc(Lf[d], d + 1, False) for d in range(4)],
                                 *[mk_kc(Lt[d], d + 1, True) for d in range(4)])
    sd = {k[len("param_"):]: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k.startswith("param_")}
    layer.load_state_dict(sd, strict=True)
    layer = layer.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    data = types.SimpleNamespace(x=x, edge_index=torch.from_numpy(g["edge_index"]).to(DEV),
                                 edge_attr=torch.from_numpy(g["edge_attr"]).to(DEV), p=torch.from_numpy(g["p"]).to(DEV))
    for k, v in g.items():
        if k.startswith("bk_"):
            t = torch.from_numpy(np.asarray(v))
            setattr(data, k[3:], (t.long() if "index" in k else t.float()).to(DEV))
    sc = layer(is_last_layer=True, data=data, save_score=False)
    assert sc.shape == g["sc"].shape
    assert rel_err(sc.detach().cpu(), g["sc"]) < TOL
    (sc * torch.from_numpy(g["wout"]).to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), g["grad_x"]) < TOL
    for d in range(4):
        for p in layer.fixed_kernelconv_set[d].parameters():
            assert p.grad is None
        kc = layer.trainable_kernelconv_set[d]
        for nme in ["x_center", "x_support", "edge_attr_support"]:
            assert rel_err(getattr(kc, nme).grad.cpu(), g[f"grad_trainable_kernelconv_set.{d}.{nme}"]) < TOL, (d, nme)
        trip = ["support_attr_sc_weight", "center_attr_sc_weight", "edge_attr_support_sc_weight"]
        refs = np.array([float(g[f"grad_trainable_kernelconv_set.{d}.{t}"]) for t in trip])
        got = np.array([getattr(kc, t).grad.item() for t in trip])
        assert np.abs(got - refs).max() <= 1e-4 * max(np.abs(refs).max(), 1e-6), (d, got, refs)


@pytest.mark.parametrize("x_dim,L1,LN", [
    (130, (10, 20, 30, 50), (10, 20, 30, 50)),      # few blocks, one column half (Fk = 160); second layer on the base tile kernels
    (300, (10, 20, 30, 50), (10, 20, 30, 50)),      # two column halves of 160 columns
    (500, (3, 5, 7, 9), (4, 6, 8, 10)),             # Fk = 512: the accumulator fills tensor memory
    (28, (60, 90, 64, 52), (10, 20, 30, 50)),       # 9 uneven blocks (18/17/17, 32/32, 30/30/30, 60), then F = 266 -> halves 160 + 128
])
def test_wide_kernels_feature_widths_and_block_shapes(x_dim, L1, LN, fwd_path):
    """The streamed-operand (wide) kernels on shapes other than configs[2]: feature widths with one and two column halves, a full
    512-column accumulator, uneven kernel blocks (blocked tile order of the coefficients with a remainder)."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth, functional
    b = synth.make_batch(20, seed=61)
    rng = np.random.default_rng(61)
    b["x"] = rng.standard_normal((b["x"].shape[0], x_dim)).astype(np.float32)
    torch.manual_seed(61)
    net = mk.MolGCN(2, *L1, *LN, x_dim=x_dim, p_dim=3, edge_attr_dim=7)
    wout = torch.randn(b["x"].shape[0], sum(LN))
    h_ref, gx_ref, params_ref, auxs = _oracle_run(net, b, wout)
    net = net.to(DEV)
    d = _to_dev(b)
    x = d["x"].clone().requires_grad_(True)
    forced = [compact_from_kernel_major([None if a is None else a["argmax"] for a in aux], DEV) for aux in auxs]
    pc0 = functional.path_counts()
    h = net(x=x, edge_index=d["edge_index"], edge_attr=d["edge_attr"], p=d["p"], save_score=False, argmax_in=forced)
    assert rel_err(h.detach().cpu(), h_ref) < TOL
    (h * wout.to(DEV)).sum().backward()
    pc1 = functional.path_counts()
    if fwd_path in ("fused", "tile"):
        assert pc1["fwd_tile"] - pc0["fwd_tile"] == 2 and pc1["bwd_tile"] - pc0["bwd_tile"] == 2, (pc0, pc1)
    assert rel_err(x.grad.cpu(), gx_ref) < TOL
    _check_param_grads(net, lambda li, dg, n: params_ref[li][dg][n].grad)
