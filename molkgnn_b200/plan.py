"""Degree-bucket plan of a collated batch, built on the GPU by the CUDA bucket pass.

Host-side mirror of what the reference precomputes per molecule with ``ToXAndPAndEdgeAttrForDeg``
(/root/reference/wrapper.py:559-672) and PyG then collates: the five per-degree attributes
``selected_index_deg*``, ``nei_index_deg*``, ``p_focal_deg*``, ``nei_p_deg*``, ``nei_edge_attr_deg*``.
Torch is only the allocator here; all work happens behind the C-ABI (include/molkgnn_b200.h).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ptr, stream_ptr, check

EDGE_PAD = 8
DEG_KEYS = ("p_focal", "nei_p", "nei_edge_attr", "selected_index", "nei_index")


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.MolKGNNError(f"molkgnn_b200: '{name}' must be a CUDA tensor (there is no CPU fallback)")


class BucketPlan(object):
    """Owns the device arrays of ``molkgnn_plan_t``; ``self.c`` is the ctypes struct handed to the kernels."""

    def __init__(self, N, E, device):
        self.N, self.E, self.device = int(N), int(E), device
        i32 = dict(dtype=torch.int32, device=device)
        self.deg = torch.empty(self.N, **i32)
        self.pos = torch.empty(self.N, **i32)
        self.sel = torch.empty(self.N, **i32)
        self.nei = torch.empty(max(self.E, 1), **i32)
        self.nei_eid = torch.empty(max(self.E, 1), **i32)
        self.ehat = torch.empty(max(self.E, 1), EDGE_PAD, dtype=torch.float32, device=device)
        self.tsign = torch.zeros(max(self.N, 1), dtype=torch.int8, device=device)
        self.in_cnt = torch.empty(self.N, **i32)
        self.in_src = torch.empty(self.N, 4, **i32)
        self.in_j = torch.empty(self.N, 4, **i32)
        self.tile_start = torch.empty(self.N // 32 + 4, **i32)      # molecule tiles (filled by molkgnn_bucket_build)
        self.tile_meta = torch.empty((self.N // 32 + 4) * int(_lib.lib().molkgnn_tile_meta_bytes()), dtype=torch.uint8,
                                     device=device)
        self.ehat_node = torch.empty(max(self.E, 1), EDGE_PAD, dtype=torch.float32, device=device)
        self.node_tile = torch.empty(max(self.N, 1), **i32)
        c = _lib.Plan()
        c.N, c.E = self.N, self.E
        for name in ("deg", "pos", "sel", "nei", "nei_eid", "ehat", "tsign", "in_cnt", "in_src", "in_j", "tile_start", "tile_meta",
                     "ehat_node", "node_tile"):
            setattr(c, name, getattr(self, name).data_ptr())
        self.c = c
        self.n = [0, 0, 0, 0]
        self.n_tiles = 0

    # ---- constructors -------------------------------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index, p, edge_attr, num_nodes):
        """One GPU CSR->degree-bucket pass over the collated ``edge_index`` (replaces wrapper.py:637-672)."""
        for t, nme in ((edge_index, "edge_index"), (p, "p"), (edge_attr, "edge_attr")):
            _require_cuda(t, nme)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise _lib.MolKGNNError("edge_index must be an int64 tensor of shape [2, E]")
        edge_index = edge_index.contiguous()
        p = p.contiguous().float()
        edge_attr = edge_attr.contiguous().float()
        E = edge_index.shape[1]
        self = cls(num_nodes, E, edge_index.device)
        L = _lib.lib()
        scratch = torch.empty(int(L.molkgnn_bucket_scratch_bytes(self.N, E)), dtype=torch.uint8, device=self.device)
        from .functional import _timed
        with _timed("bucket_build"):
            check(L.molkgnn_bucket_build(C.byref(self.c), ptr(edge_index), ptr(p), p.shape[1], ptr(edge_attr),
                                         edge_attr.shape[1], ptr(scratch), stream_ptr()))
        self.n = list(self.c.n)
        self.n_tiles = int(self.c.n_tiles)          # 0: no molecule tiling (bucket-order kernels are used)
        self._keep = (edge_index, p, edge_attr)
        return self

    @classmethod
    def from_reference_tensors(cls, num_nodes, selected_index, nei_index, p_focal, nei_p, nei_edge_attr):
        """Plan from the per-degree attributes a reference batch already carries (kernels.py:628-645); each argument is
        a list of four tensors (degree 1..4), possibly empty."""
        dev = None
        for t in selected_index:
            if t is not None and t.numel():
                _require_cuda(t, "selected_index")
                dev = t.device
        if dev is None:
            raise _lib.MolKGNNError("all degree buckets are empty")
        n = [int(t.numel()) if t is not None else 0 for t in selected_index]
        E = sum(n[d] * (d + 1) for d in range(4))
        self = cls(num_nodes, E, dev)
        Fe, p_dim = None, 3
        keep = []

        def arr(ts, dtype):
            a = (C.c_void_p * 4)()
            for d in range(4):
                t = ts[d]
                if n[d] == 0 or t is None:
                    a[d] = None
                    continue
                t = t.to(device=dev, dtype=dtype).contiguous()
                keep.append(t)
                a[d] = t.data_ptr()
            return a

        for d in range(4):
            if n[d]:
                Fe = int(nei_edge_attr[d].shape[-1])
                p_dim = int(p_focal[d].shape[-1])
        for d in range(4):
            self.c.n[d] = n[d]
        check(_lib.lib().molkgnn_plan_from_buckets(C.byref(self.c), arr(selected_index, torch.int64),
                                                   arr(nei_index, torch.int64), arr(p_focal, torch.float32),
                                                   arr(nei_p, torch.float32), p_dim,
                                                   arr(nei_edge_attr, torch.float32), Fe, stream_ptr()))
        self.n = n
        self._keep = keep
        return self

    # ---- views --------------------------------------------------------------------------------------------
    @property
    def boff(self):
        return list(self.c.boff)

    @property
    def eoff(self):
        return list(self.c.eoff)

    def scoff(self, Ls):
        """float offsets of the per-degree compact score blocks for kernel counts ``Ls`` + total size"""
        off, tot = [], 0
        for d in range(4):
            off.append(tot)
            tot += self.n[d] * int(Ls[d])
        return off, tot

    def export(self, d, p, edge_attr):
        """Reference-format attributes of degree ``d`` (int64 indices, raw gathers) -- bit-exact with wrapper.py:595-635."""
        n = self.n[d - 1]
        dev = self.device
        p = p.contiguous().float()
        edge_attr = edge_attr.contiguous().float()
        pd, Fe = p.shape[1], edge_attr.shape[1]
        out = dict(selected_index=torch.empty(n, dtype=torch.int64, device=dev),
                   nei_index=torch.empty(n * d, dtype=torch.int64, device=dev),
                   p_focal=torch.empty(n, pd, dtype=torch.float32, device=dev),
                   nei_p=torch.empty(n, d, pd, dtype=torch.float32, device=dev),
                   nei_edge_attr=torch.empty(n, d, Fe, dtype=torch.float32, device=dev))
        check(_lib.lib().molkgnn_bucket_export(C.byref(self.c), d, ptr(p), pd, ptr(edge_attr), Fe,
                                               ptr(out["selected_index"]), ptr(out["nei_index"]), ptr(out["p_focal"]),
                                               ptr(out["nei_p"]), ptr(out["nei_edge_attr"]), stream_ptr()))
        return out


class ToXAndPAndEdgeAttrForDeg(object):
    """Drop-in for the reference pre-transform (wrapper.py:559-672): attaches the 20 per-degree attributes to ``data``.
    Works on a single graph or on an already collated batch (results equal per-graph transform + PyG collation).
    Empty buckets give empty tensors like the reference (wrapper.py:627-630)."""

    def __call__(self, data):
        dev_in = data.x.device
        cuda = torch.device("cuda", torch.cuda.current_device()) if not data.x.is_cuda else data.x.device
        ei, p, ea = data.edge_index.to(cuda), data.p.to(cuda), data.edge_attr.to(cuda)
        plan = BucketPlan.from_edge_index(ei, p, ea, data.x.shape[0])
        for d in range(1, 5):
            if plan.n[d - 1] == 0:
                out = dict(p_focal=torch.empty(0, p.shape[1]), nei_p=torch.Tensor(), nei_edge_attr=torch.Tensor(),
                           selected_index=torch.empty(0, dtype=torch.int64),
                           nei_index=torch.empty(0, dtype=torch.int64))
            else:
                out = plan.export(d, p, ea)
            for k in DEG_KEYS:
                setattr(data, f"{k}_deg{d}", out[k].to(dev_in))
        return data
