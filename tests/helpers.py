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


# ---- GPU-side helpers -------------------------------------------------------------------------------------------
def compact_from_kernel_major(per_degree, device):
    """[argmax [L_d,n_d] or None x4] -> uint8 compact vector (degree blocks, row-major [n_d, L_d])."""
    parts = []
    for a in per_degree:
        if a is None:
            continue
        parts.append(torch.as_tensor(a).t().contiguous().reshape(-1).to(torch.uint8))
    if not parts:
        return torch.zeros(1, dtype=torch.uint8, device=device)
    return torch.cat(parts).to(device)


def kernel_major_from_compact(vec, n, Ls):
    """inverse of compact_from_kernel_major: -> list of 4 tensors [L_d, n_d] (None for empty buckets)."""
    out, off = [], 0
    vec = vec.cpu()
    for d in range(4):
        cnt = n[d] * Ls[d]
        out.append(vec[off:off + cnt].reshape(n[d], Ls[d]).t().contiguous() if cnt else None)
        off += cnt
    return out


def module_from_golden(g, device):
    import molkgnn_b200 as mk
    L1, LN = [int(v) for v in g["L1"]], [int(v) for v in g["LN"]]
    net = mk.MolGCN(num_layers=int(g["num_layers"]), num_kernel1_1hop=L1[0], num_kernel2_1hop=L1[1],
                    num_kernel3_1hop=L1[2], num_kernel4_1hop=L1[3], num_kernel1_Nhop=LN[0], num_kernel2_Nhop=LN[1],
                    num_kernel3_Nhop=LN[2], num_kernel4_Nhop=LN[3], x_dim=g["x"].shape[1], p_dim=3,
                    edge_attr_dim=g["edge_attr"].shape[1])
    sd = {k[len("param_"):]: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k.startswith("param_")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    return net.to(device)


def params_from_module(net, dtype=torch.float32, requires_grad=False):
    """MolGCN module -> oracle parameter structure (CPU)."""
    out = []
    for layer in net.layers:
        lp = []
        for kc in layer.trainable_kernelconv_set:
            lp.append({n: getattr(kc, n).detach().cpu().to(dtype).clone().requires_grad_(requires_grad)
                       for n in PARAM_NAMES})
        out.append(lp)
    return out


def tile_schedule_model(nn, G):
    """Plain-numpy model of k_tile_order (csrc/bucket.cu; no reference counterpart -- the tile schedule is internal to the
    tile-major backward kernels).  nn[t] = nodes of tile t, G persistent CTAs.  Returns `order`: CTA c walks
    order[beg_c : beg_c + cnt_c] with cnt_c = q + (c < r), n = q G + r (csrc/tile.cuh TileWalk).  Tiles are sorted by node
    count (stable); the r long CTAs take the (q + 1) r smallest, the others the rest, dealt boustrophedon in each group."""
    nn = np.asarray(nn)
    n = len(nn)
    q, r = divmod(n, G)
    srt = np.argsort(nn, kind="stable")
    order = np.full(n, -1, dtype=np.int64)
    nlong, Gs = (q + 1) * r, G - r
    for s, t in enumerate(srt):
        if s < nlong:
            k, i = divmod(s, r)
            c = r - 1 - i if k & 1 else i
            pos = c * (q + 1) + k
        else:
            k, i = divmod(s - nlong, Gs)
            c = Gs - 1 - i if k & 1 else i
            pos = nlong + c * q + k
        order[pos] = t
    return order


def oracle_chunked(net, b, wout, chunk=32):
    """The CPU oracle over a LARGE batch in chunks of whole molecules (molecules are independent graphs, so h / grad_x are
    per-chunk and kernel-parameter gradients add over chunks -- tests/test_fullsize_gpu.py proves both for the CUDA path).

    b: numpy batch of synth.collate (needs "ptr"; edges grouped molecule by molecule); wout [N, K] dL/dh.
    Returns dict(h, grad_x, grads {(li, d, name): tensor}, argmax [layer][d-1] -> [L_d, n_d] int64 (or None),
    S [layer][d-1] -> [L_d, P_d, n_d] fp32 (or None)) in FULL-batch bucket order (chunks are contiguous node ranges and
    bucket rows ascend with the node id, so per-degree concatenation over chunks is the full batch's order)."""
    params = params_from_module(net, requires_grad=True)
    ptr = np.asarray(b["ptr"])
    src = b["edge_index"][0]
    nl = len(params)
    hs, gxs = [], []
    am = [[[] for _ in range(4)] for _ in range(nl)]
    Ss = [[[] for _ in range(4)] for _ in range(nl)]
    for m0 in range(0, len(ptr) - 1, chunk):
        m1 = min(m0 + chunk, len(ptr) - 1)
        n0, n1 = int(ptr[m0]), int(ptr[m1])
        e0, e1 = int(np.searchsorted(src, n0, side="left")), int(np.searchsorted(src, n1, side="left"))
        ei = b["edge_index"][:, e0:e1] - n0
        assert ei.min() >= 0 and ei.max() < n1 - n0, "edges must be grouped molecule by molecule"
        bk = orc.buckets_to_torch(orc.bucket_pass(ei, n1 - n0, b["p"][n0:n1], b["edge_attr"][e0:e1]))
        x = torch.from_numpy(b["x"][n0:n1]).clone().requires_grad_(True)
        h, auxs = orc.molgcn_forward(params, x, torch.from_numpy(ei), bk, return_aux=True)
        (h * wout[n0:n1]).sum().backward()          # parameter .grad accumulates over the chunks
        hs.append(h.detach())
        gxs.append(x.grad)
        for li in range(nl):
            for d in range(4):
                if auxs[li][d] is not None:
                    am[li][d].append(auxs[li][d]["argmax"].detach())
                    Ss[li][d].append(auxs[li][d]["S"].detach().float())
    cat = lambda lst, dim: torch.cat(lst, dim=dim) if lst else None  # noqa: E731
    grads = {(li, d, n): params[li][d][n].grad for li in range(nl) for d in range(4) for n in PARAM_NAMES
             if params[li][d][n].grad is not None}
    return dict(h=torch.cat(hs), grad_x=torch.cat(gxs), grads=grads,
                argmax=[[cat(am[li][d], 1) for d in range(4)] for li in range(nl)],
                S=[[cat(Ss[li][d], 2) for d in range(4)] for li in range(nl)])


def elementwise_close(a, b, rtol=1e-5, floor_frac=1e-5):
    """Element-wise check beside the max-normalised one: |a - b| <= rtol |b| + floor_frac * max|b|.  The absolute floor is the
    rounding of an fp32 sum whose terms are O(max|b|) (entries that cancel to ~0 cannot be relatively exact).
    -> (ok, worst excess ratio)"""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    bound = rtol * b.abs() + floor_frac * float(b.abs().max().clamp_min(1e-30))
    ratio = float(((a - b).abs() / bound).max())
    return ratio <= 1.0, ratio
