"""Direct parameter gradients (functional.MolGCNFn direct mode: the backward writes p.grad itself) against the autograd path
(72 gradient tensors returned to the engine): same bits, accumulation, persistent buffer, hooks fall back to autograd."""
import numpy as np
import pytest
import torch

import molkgnn_b200 as mk
from molkgnn_b200 import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def _setup(n_mol=96, seed=3):
    b = synth.make_batch(n_mol, seed=seed)
    t = {k: torch.from_numpy(b[k]).to(DEV) for k in ("x", "p", "edge_index", "edge_attr")}
    torch.manual_seed(1)
    net = mk.MolGCN(3, 10, 20, 30, 50, 10, 20, 30, 50, x_dim=28, p_dim=3, edge_attr_dim=7).to(DEV)
    wout = torch.randn(t["x"].shape[0], 110, device=DEV)
    return net, t, wout


def _run(net, t, wout):
    x = t["x"].detach().requires_grad_(True)
    h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    (h * wout).sum().backward()
    return h.detach().clone(), x.grad.clone()


def _grads(net):
    return {n: (None if p.grad is None else p.grad.clone()) for n, p in net.named_parameters()}


def test_direct_equals_autograd_bitwise():
    net, t, wout = _setup()
    assert net.__dict__.get("direct_param_grads") in (None, True)
    h_d, gx_d = _run(net, t, wout)
    g_d = _grads(net)
    net.zero_grad(set_to_none=True)
    net.direct_param_grads = False
    h_a, gx_a = _run(net, t, wout)
    g_a = _grads(net)
    assert torch.equal(h_d, h_a) and torch.equal(gx_d, gx_a)
    assert set(g_d) == set(g_a)
    n_with = 0
    for n in g_a:
        assert (g_d[n] is None) == (g_a[n] is None), n
        if g_a[n] is not None:
            assert torch.equal(g_d[n], g_a[n]), n
            n_with += 1
    assert n_with >= 3 * 4 * 6


def test_accumulation_and_persistent_buffer():
    net, t, wout = _setup()
    _run(net, t, wout)
    g1 = _grads(net)
    p0 = next(p for p in net.parameters() if p.grad is not None)
    ptr1 = p0.grad.data_ptr()
    _run(net, t, wout)                                   # .grad still set: the second backward must ADD
    for n, p in net.named_parameters():
        if g1[n] is not None:
            assert torch.equal(p.grad, 2 * g1[n]), n
    net.zero_grad(set_to_none=True)
    _run(net, t, wout)                                   # all None again: back in the persistent buffer, same values as step 1
    assert p0.grad.data_ptr() == ptr1
    for n, p in net.named_parameters():
        if g1[n] is not None:
            assert torch.equal(p.grad, g1[n]), n


def test_frozen_parameters_and_inference():
    net, t, wout = _setup()
    frozen = [p for n, p in net.named_parameters() if n.startswith("layers.1.")]
    for p in frozen:
        p.requires_grad_(False)
    _run(net, t, wout)
    assert all(p.grad is None for p in frozen)
    assert any(p.grad is not None for n, p in net.named_parameters() if n.startswith("layers.0."))
    with torch.no_grad():
        h = net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    assert not h.requires_grad
    for p in net.parameters():
        p.requires_grad_(False)
    h = net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    assert not h.requires_grad


def test_hooks_and_autograd_grad_use_the_autograd_edges():
    net, t, wout = _setup()
    seen = []
    p0 = net.layers[0].trainable_kernelconv_set[2].x_center
    p0.register_hook(lambda g: seen.append(float(g.abs().sum())))
    _run(net, t, wout)
    assert len(seen) == 1 and seen[0] > 0                 # a hooked parameter switches the module to the autograd path
    net2, t2, wout2 = _setup()
    net2.direct_param_grads = False
    x = t2["x"].detach().requires_grad_(True)
    h = net2(x=x, edge_index=t2["edge_index"], edge_attr=t2["edge_attr"], p=t2["p"], save_score=False)
    q0 = net2.layers[0].trainable_kernelconv_set[2].x_center
    (g,) = torch.autograd.grad((h * wout2).sum(), [q0])
    assert g.shape == q0.shape and float(g.abs().sum()) > 0 and q0.grad is None


def test_parameters_get_gradients_when_x_needs_none():
    """x without requires_grad: the graph reaches the backward through the anchor tensor alone."""
    net, t, wout = _setup()
    h = net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    assert h.requires_grad
    (h * wout).sum().backward()
    g_direct = _grads(net)
    net.zero_grad(set_to_none=True)
    net.direct_param_grads = False
    h = net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
    (h * wout).sum().backward()
    n_with = 0
    for n, p in net.named_parameters():
        assert (p.grad is None) == (g_direct[n] is None), n
        if p.grad is not None:
            assert torch.equal(p.grad, g_direct[n]), n
            n_with += 1
    assert n_with >= 3 * 4 * 6
