"""CPU ORACLE for the MolKGNN molecular-kernel convolution.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference hot path so that the CUDA product
in ``molkgnn_b200/`` can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product never does.

Parity pin: the restatement is validated against the UNMODIFIED reference modules
(/root/reference/models/MolKGNN/{kernels,KernelLayer}.py and wrapper.ToXAndPAndEdgeAttrForDeg), imported
in the build container through the test-only stub in ``tests/stubs`` by ``tools/make_golden.py``; the
reference's outputs are committed as fixtures under ``tests/golden/`` and ``tests/test_oracle.py`` checks
this file against them (plus the reference's only in-tree known-answer vector, kernels.py:161-170).
Third-party arithmetic that is NOT under /root/reference (torch ``cosine_similarity``/``max``/``mean``,
PyG ``propagate``/``Data.__inc__``) is restated from its documented behaviour: parity unpinned at those
two library boundaries; everything the reference itself computes is pinned.

Function -> reference map
  perm_table                      kernels.py:89-130        (KernelConv.permute tables)
  cosine_mean                     kernels.py:154-195       (calculate_average_similarity_score)
  kernel_conv_forward             kernels.py:353-425       (calculate_total_score) incl. :230-275, :197-225
  chirality_sign                  kernels.py:279-350       (get_chirality_sign), vectorised
  kernel_set_conv_forward         kernels.py:610-751       (BaseKernelSetConv.forward)
  propagate                       KernelLayer.py:119-123   (MessagePassing aggr='add', message = sim_sc_j)
  molgcn_forward                  KernelLayer.py:107-120
  bucket_pass                     wrapper.py:567-635       (ToXAndPAndEdgeAttrForDeg) + PyG collate semantics
"""
from __future__ import annotations

from itertools import permutations

import numpy as np
import torch

EPS = 1e-8  # torch.nn.CosineSimilarity default eps (kernels.py:189)

_D4_EVEN = [(0, 1, 2, 3), (0, 2, 3, 1), (0, 3, 1, 2), (1, 0, 3, 2), (1, 2, 0, 3), (1, 3, 2, 0),
            (2, 0, 1, 3), (2, 1, 3, 0), (2, 3, 0, 1), (3, 0, 2, 1), (3, 1, 0, 2), (3, 2, 1, 0)]


def perm_table(d: int):
    """kernels.py:109-128: all d! permutations in lexicographic order for d != 4, the 12 listed even ones for d == 4."""
    if d != 4:
        return [tuple(p) for p in permutations(range(d))]
    return list(_D4_EVEN)


def _normalize(t: torch.Tensor) -> torch.Tensor:
    # torch >= 2.0 cosine_similarity: (x / max(||x||, eps)) . (y / max(||y||, eps))
    return t / torch.linalg.vector_norm(t, 2, dim=-1, keepdim=True).clamp_min(EPS)


def cosine_mean(t1, t2, avg=True):
    """kernels.py:154-195 with sim_dim=-1, avg_dim=-2."""
    sc = (_normalize(t1) * _normalize(t2)).sum(-1)
    return sc.mean(-1) if avg else sc


def chirality_sign(p_nei, x_nei, best_p_support):
    """kernels.py:279-350, vectorised.  p_nei [n,4,3] (already calibrated), x_nei [n,4,F], best_p_support [L,n,4,3] -> int64 [L,n]."""
    n = x_nei.shape[0]
    L = best_p_support.shape[0]
    dup = torch.zeros(n, dtype=torch.bool)
    for a in range(4):
        for b in range(a + 1, 4):
            dup |= (x_nei[:, a] == x_nei[:, b]).all(-1)  # torch.equal on the two rows (kernels.py:314)

    def triple(t):  # sign(t2 . (t0 x t1)), only neighbours 0..2 are used (kernels.py:327-341)
        c = torch.cross(t[..., 0, :], t[..., 1, :], dim=-1)
        return torch.sign((t[..., 2, :] * c).sum(-1))

    s_nei = triple(p_nei)            # [n]
    s_sup = triple(best_p_support)   # [L,n]
    chi = torch.where(s_nei.unsqueeze(0) == s_sup, 1, -1).to(torch.int64)
    chi[:, dup] = 1
    assert chi.shape == (L, n)
    return chi


def kernel_conv_forward(params, x_focal, p_focal, x_neighbor, p_neighbor, edge_attr_neighbor,
                        is_last_layer=False, return_aux=False, force_argmax=None):
    """One degree bucket (kernels.py:353-425).  ``params``: dict with x_center [L,F], x_support [L,d,F],
    edge_attr_support [L,d,Fe], p_support [L,d,3], support_attr_sc_weight, center_attr_sc_weight,
    edge_attr_support_sc_weight (0-dim).  Returns sc [L,n] (kernel-major, like the reference).

    ``force_argmax`` [L,n] (optional, parity harness only): use this permutation index instead of the arg-max
    for everything downstream of kernels.py:373 (teacher forcing across structurally tied permutations); the
    free-running arg-max is still reported in aux['argmax']."""
    xs, xc, es, ps = params["x_support"], params["x_center"], params["edge_attr_support"], params["p_support"]
    L, d, F = xs.shape
    n = x_focal.shape[0]
    perms = torch.tensor(perm_table(d), dtype=torch.long)              # [P,d]
    p_neighbor = p_neighbor - p_focal.unsqueeze(1)                     # :356

    xn = _normalize(x_neighbor)                                        # [n,d,F]
    sn = _normalize(xs)                                                # [L,d,F]
    A = torch.einsum("njf,ksf->knjs", xn, sn)                          # [L,n,d(j),d(s)]
    # S[k,pi,n] = mean_j A[k,n,j,perm_pi[j]]  (sequential sum over j, then true division by d; :194)
    idx = perms.t().reshape(1, 1, d, -1).expand(L, n, d, perms.shape[0])   # [L,n,j,P] -> s index
    Aj = torch.gather(A, 3, idx)                                       # [L,n,j,P]
    S = Aj[:, :, 0, :]
    for j in range(1, d):
        S = S + Aj[:, :, j, :]
    S = (S / d).permute(0, 2, 1)                                       # [L,P,n]
    best_S, best_i = torch.max(S, dim=1)                               # first max wins; [L,n]
    free_i = best_i
    if force_argmax is not None:
        best_i = force_argmax.long()
        best_S = torch.gather(S, 1, best_i.unsqueeze(1)).squeeze(1)

    C = torch.einsum("nf,kf->kn", _normalize(x_focal), _normalize(xc))  # :254-270

    best_perm = perms[best_i]                                          # [L,n,d]
    en = _normalize(edge_attr_neighbor)                                # [n,d,Fe]
    esn = _normalize(es)                                               # [L,d,Fe]
    Ee = torch.einsum("nje,kse->knjs", en, esn)                        # [L,n,j,s]
    Ej = torch.gather(Ee, 3, best_perm.unsqueeze(-1)).squeeze(-1)      # [L,n,j]
    E = Ej[:, :, 0]
    for j in range(1, d):
        E = E + Ej[:, :, j]
    E = E / d

    chi = None
    if d == 4 and is_last_layer:                                       # :396
        best_p_support = ps[torch.arange(L).view(L, 1, 1), best_perm]  # [L,n,4,3]  (:197-225)
        chi = chirality_sign(p_neighbor, x_neighbor, best_p_support)

    es_, ec_, ee_ = (torch.exp(params["support_attr_sc_weight"]), torch.exp(params["center_attr_sc_weight"]),
                     torch.exp(params["edge_attr_support_sc_weight"]))
    den = es_ + ec_ + ee_
    ws, wc, we = es_ / den, ec_ / den, ee_ / den
    sc = (best_S * ws + C * wc + E * we) / (ws + wc + we)
    if chi is not None:
        sc = sc * chi
    if return_aux:
        return sc, dict(argmax=free_i, used=best_i, S=S, best_S=best_S, C=C, E=E, chi=chi)
    return sc


def kernel_set_conv_forward(layer_params, x, buckets, is_last_layer=False, return_aux=False, force_argmax=None):
    """BaseKernelSetConv.forward (kernels.py:610-751) -> sc [N,K].  ``layer_params``: list of 4 per-degree param
    dicts; ``buckets``: output of :func:`bucket_pass` (torch tensors)."""
    N = x.shape[0]
    Ls = [lp["x_center"].shape[0] for lp in layer_params]
    K = sum(Ls)
    rows, aux_all = [], []
    sc_full = torch.zeros(K, N, dtype=x.dtype)
    cols = []
    r0 = 0
    for d in range(1, 5):
        b = buckets[d]
        sel, nei = b["selected_index"], b["nei_index"]
        if sel.numel() == 0:
            r0 += Ls[d - 1]
            aux_all.append(None)
            continue
        x_focal = x.index_select(0, sel)
        x_nei = x.index_select(0, nei).reshape(-1, d, x.shape[1])
        out = kernel_conv_forward(layer_params[d - 1], x_focal, b["p_focal"].to(x.dtype), x_nei,
                                  b["nei_p"].to(x.dtype), b["nei_edge_attr"].to(x.dtype), is_last_layer,
                                  return_aux=return_aux,
                                  force_argmax=None if force_argmax is None else force_argmax[d - 1])
        if return_aux:
            out, aux = out
            aux_all.append(aux)
        blk = torch.zeros(K, sel.numel(), dtype=x.dtype)
        blk[r0:r0 + Ls[d - 1]] = out
        rows.append(blk)
        cols.append(sel)
        r0 += Ls[d - 1]
    sc = torch.cat(rows, dim=1)
    order = torch.sort(torch.cat(cols), dim=0)[1]
    sc = sc[:, order].T
    if return_aux:
        return sc, aux_all
    return sc


def propagate(edge_index, sim_sc):
    """KernelLayer.py:119-123 with PyG aggr='add', flow source_to_target: h[i] = sum_{e: dst(e)=i} sim_sc[src(e)], edge order."""
    out = torch.zeros_like(sim_sc)
    return out.index_add(0, edge_index[1], sim_sc.index_select(0, edge_index[0]))


def molgcn_forward(all_params, x, edge_index, buckets, return_aux=False, force_argmax=None):
    """MolGCN.forward (KernelLayer.py:107-120): h <- propagate(KernelSetConv_i(h)) for each layer; last layer flagged."""
    h = x
    auxs = []
    for i, lp in enumerate(all_params):
        out = kernel_set_conv_forward(lp, h, buckets, is_last_layer=(i == len(all_params) - 1), return_aux=return_aux,
                                      force_argmax=None if force_argmax is None else force_argmax[i])
        if return_aux:
            out, aux = out
            auxs.append(aux)
        h = propagate(edge_index, out)
    return (h, auxs) if return_aux else h


# ----------------------------------------------------------------------------------------------------------------
# degree bucketing (integer work: bit-exact bar)
# ----------------------------------------------------------------------------------------------------------------

def bucket_pass(edge_index: np.ndarray, num_nodes: int, p: np.ndarray, edge_attr: np.ndarray):
    """wrapper.py:567-635 applied to a collated batch (equals per-molecule transform + PyG collation because
    ``selected_index``/``nei_index`` carry 'index' in their key -> offset by the cumulative node count).

    Returns {d: dict(selected_index i64[n_d], nei_index i64[n_d*d], p_focal f32[n_d,3], nei_p f32[n_d,d,3],
    nei_edge_attr f32[n_d,d,Fe])} for d = 1..4 and 'deg' (int64 [N]).  Neighbours of a focal node are listed in
    edge order (wrapper.py:567-572); bond attributes are taken from row 2*(eid//2) (wrapper.py:586-591)."""
    src, dst = edge_index[0], edge_index[1]
    deg = np.bincount(src, minlength=num_nodes).astype(np.int64)
    order = np.argsort(src, kind="stable")          # edges grouped by source, edge order kept inside a group
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    out = {"deg": deg}
    for d in range(1, 5):
        sel = np.nonzero(deg == d)[0].astype(np.int64)
        if sel.size:
            eids = order[(rowptr[sel][:, None] + np.arange(d)[None, :])]       # [n_d,d]
            nei = dst[eids]
            nei_p = p[nei]
            nei_ea = edge_attr[2 * (eids // 2)]
        else:
            nei = np.zeros((0, d), dtype=np.int64)
            nei_p = np.zeros((0, d, p.shape[1]), dtype=p.dtype)
            nei_ea = np.zeros((0, d, edge_attr.shape[1]), dtype=edge_attr.dtype)
        out[d] = dict(selected_index=sel, nei_index=nei.reshape(-1).astype(np.int64), p_focal=p[sel],
                      nei_p=nei_p, nei_edge_attr=nei_ea)
    return out


def buckets_to_torch(b):
    return {k: ({kk: torch.from_numpy(np.ascontiguousarray(vv)) for kk, vv in v.items()} if isinstance(v, dict)
                else torch.from_numpy(v)) for k, v in b.items()}


# ----------------------------------------------------------------------------------------------------------------
# parameter construction in the reference's RNG order (kernels.py:50-53, 762-774; KernelLayer.py:27-46)
# ----------------------------------------------------------------------------------------------------------------

def init_layer_params(Ls, D, node_attr_dim, edge_attr_dim, dtype=torch.float32, requires_grad=False):
    out = []
    for d, L in enumerate(Ls, start=1):
        prm = dict(x_center=torch.randn(L, node_attr_dim), x_support=torch.randn(L, d, node_attr_dim),
                   edge_attr_support=torch.randn(L, d, edge_attr_dim), p_support=torch.randn(L, d, D),
                   length_sc_weight=torch.tensor(0.2), angle_sc_weight=torch.tensor(0.2),
                   center_attr_sc_weight=torch.tensor(0.2), support_attr_sc_weight=torch.tensor(0.2),
                   edge_attr_support_sc_weight=torch.tensor(0.2))
        prm = {k: v.to(dtype).requires_grad_(requires_grad) for k, v in prm.items()}
        out.append(prm)
    return out


def init_molgcn_params(num_layers, L_1hop, L_Nhop, x_dim, p_dim=3, edge_attr_dim=7, **kw):
    layers = [init_layer_params(L_1hop, p_dim, x_dim, edge_attr_dim, **kw)]
    f = sum(L_1hop)
    for _ in range(num_layers - 1):
        layers.append(init_layer_params(L_Nhop, p_dim, f, edge_attr_dim, **kw))
        f = sum(L_Nhop)
    return layers


# ----------------------------------------------------------------------------------------------------------------
# batch assembly (integer / byte work: bit-exact bar) -- PyG 2.0.x `Batch.from_data_list` as the reference's DataLoaders
# use it (data.py:136-229): every attribute concatenated along its __cat_dim__ (dim 0; -1 for keys containing "index"),
# keys containing "index" incremented by the number of nodes in front of the graph (__inc__), plus `batch` and `ptr`.
# torch_geometric is not under /root/reference: restated from its documented behaviour (parity unpinned at that boundary).
# ----------------------------------------------------------------------------------------------------------------
def collate_pyg(mols):
    """mols: sequence of objects with x [n,F], p [n,P], edge_index [2,e] (local ids), edge_attr [e,Fe] -> dict of numpy arrays"""
    n = np.array([np.asarray(m.x).shape[0] for m in mols], dtype=np.int64)
    ptr = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    return dict(x=np.concatenate([np.asarray(m.x) for m in mols], axis=0),
                p=np.concatenate([np.asarray(m.p) for m in mols], axis=0),
                edge_attr=np.concatenate([np.asarray(m.edge_attr) for m in mols], axis=0),
                edge_index=np.concatenate([np.asarray(m.edge_index) + off for m, off in zip(mols, ptr[:-1])], axis=1),
                batch=np.repeat(np.arange(len(mols), dtype=np.int64), n), ptr=ptr)
