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

    # device arrays of molkgnn_plan_t: (name, dtype, elements as a function of N, E)
    _ARRAYS = (("deg", torch.int32, lambda N, E: N), ("pos", torch.int32, lambda N, E: N),
               ("sel", torch.int32, lambda N, E: N), ("nei", torch.int32, lambda N, E: max(E, 1)),
               ("nei_eid", torch.int32, lambda N, E: max(E, 1)), ("ehat", torch.float32, lambda N, E: max(E, 1) * EDGE_PAD),
               ("tsign", torch.int8, lambda N, E: max(N, 1)), ("in_cnt", torch.int32, lambda N, E: N),
               ("in_src", torch.int32, lambda N, E: 4 * N), ("in_j", torch.int32, lambda N, E: 4 * N),
               ("tile_start", torch.int32, lambda N, E: N // 32 + 4), ("tile_meta", torch.uint8, None),
               ("ehat_node", torch.float32, lambda N, E: max(E, 1) * EDGE_PAD),
               ("node_tile", torch.int32, lambda N, E: max(N, 1)),
               ("tile_order", torch.int32, lambda N, E: N // 32 + 4))
    _SHAPES = {"ehat": (-1, EDGE_PAD), "ehat_node": (-1, EDGE_PAD), "in_src": (-1, 4), "in_j": (-1, 4)}
    _layout_cache = {}

    def __init__(self, N, E, device, scratch_bytes=0, zero_tsign=True):
        """One allocation holds every device array of the plan (+ the scratch of the bucket pass behind them)."""
        self.N, self.E, self.device = int(N), int(E), device
        key = (self.N, self.E)
        lay = BucketPlan._layout_cache.get(key)
        if lay is None:
            meta = (self.N // 32 + 4) * int(_lib.lib().molkgnn_tile_meta_bytes())
            off, lay = 0, {}
            for name, dt, fn in self._ARRAYS:
                nbytes = meta if fn is None else fn(self.N, self.E) * torch.empty(0, dtype=dt).element_size()
                lay[name] = (off, nbytes, dt)
                off += (nbytes + 127) // 128 * 128
            lay["_total"] = off
            if len(BucketPlan._layout_cache) > 64:
                BucketPlan._layout_cache.clear()
            BucketPlan._layout_cache[key] = lay
        self._lay = lay
        self._buf = torch.empty(lay["_total"] + int(scratch_bytes), dtype=torch.uint8, device=device)
        base = self._buf.data_ptr()
        c = _lib.Plan()
        c.N, c.E = self.N, self.E
        for name, _, _ in self._ARRAYS:
            setattr(c, name, base + lay[name][0])
        if zero_tsign:      # plans built from reference tensors with p_dim != 3 never write the chirality signs
            o, nb, _ = lay["tsign"]
            self._buf[o:o + nb].zero_()
        self.c = c
        self.n = [0, 0, 0, 0]
        self.n_tiles = 0

    def __getattr__(self, name):
        # tensor views of the device arrays, created on demand (tests, export)
        lay = self.__dict__.get("_lay")
        if lay is not None and name in lay and not name.startswith("_"):
            o, nb, dt = lay[name]
            t = self._buf[o:o + nb].view(dt)
            return t.view(*self._SHAPES[name]) if name in self._SHAPES else t
        raise AttributeError(name)

    @property
    def scratch_ptr(self):
        return C.c_void_p(self._buf.data_ptr() + self._lay["_total"])

    # ---- constructors -------------------------------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index, p, edge_attr, num_nodes, ref_rows=None):
        """One GPU CSR->degree-bucket pass over the collated ``edge_index`` (replaces wrapper.py:637-672)."""
        return cls.begin_from_edge_index(edge_index, p, edge_attr, num_nodes, ref_rows=ref_rows).finish()

    @staticmethod
    def ref_rows_from_kwargs(kw):
        """The raw per-degree DATA tensors of the reference's forward protocol (kernels.py:628-645), if the caller handed all of
        them: the conv must read ITS bond rows / coordinates from these (kernels.py:679, 356), not from ``edge_attr`` / ``p``
        (MolKGNNNet batch-normalises ``edge_attr`` before calling MolGCN, MolKGNNNet.py:116-119).  -> dict or None."""
        names = [f"nei_edge_attr_deg{d}" for d in range(1, 5)]
        if not all(n in kw and kw[n] is not None for n in names):
            return None
        return dict(nei_edge_attr=[kw[n] for n in names], p_focal4=kw.get("p_focal_deg4"), nei_p4=kw.get("nei_p_deg4"))

    @classmethod
    def begin_from_edge_index(cls, edge_index, p, edge_attr, num_nodes, ref_rows=None):
        """First half of the pass: the counting kernels are queued, nothing is waited for.  ``finish()`` (called by
        ``from_edge_index`` / by the MolGCN forward) does the one host round trip for the bucket sizes; whatever the caller
        queues or computes in between overlaps with the counting kernels.

        ``ref_rows`` (``ref_rows_from_kwargs``): take the bond rows and the degree-4 coordinates from the reference batch's
        own per-degree tensors instead of gathering them from ``edge_attr`` / ``p``."""
        for t, nme in ((edge_index, "edge_index"), (p, "p"), (edge_attr, "edge_attr")):
            _require_cuda(t, nme)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise _lib.MolKGNNError("edge_index must be an int64 tensor of shape [2, E]")
        edge_index = edge_index.contiguous()
        p = p.contiguous().float()
        edge_attr = edge_attr.contiguous().float()
        E = edge_index.shape[1]
        L = _lib.lib()
        self = cls(num_nodes, E, edge_index.device, scratch_bytes=int(L.molkgnn_bucket_scratch_bytes(int(num_nodes), E)),
                   zero_tsign=False)
        check(L.molkgnn_bucket_build_begin(C.byref(self.c), ptr(edge_index), ptr(p), p.shape[1], ptr(edge_attr),
                                           edge_attr.shape[1], self.scratch_ptr, stream_ptr()))
        self._keep = (edge_index, p, edge_attr)
        self._ref = None
        if ref_rows is not None:
            dev = edge_index.device
            nea, nrows = (C.c_void_p * 4)(), (C.c_int64 * 4)()
            keep = []
            for d in range(4):
                t = ref_rows["nei_edge_attr"][d]
                if t is None or t.numel() == 0:           # empty bucket: torch.Tensor() in the reference (wrapper.py:627-630)
                    nea[d], nrows[d] = None, 0
                    continue
                if t.shape[-1] != edge_attr.shape[1]:
                    raise _lib.MolKGNNError(f"nei_edge_attr_deg{d + 1} has {t.shape[-1]} columns, edge_attr {edge_attr.shape[1]}")
                t = t.to(device=dev, dtype=torch.float32).contiguous()
                keep.append(t)
                nea[d], nrows[d] = t.data_ptr(), t.numel() // t.shape[-1]
            pf4, np4 = ref_rows.get("p_focal4"), ref_rows.get("nei_p4")
            if pf4 is not None and np4 is not None and pf4.numel() and np4.numel() and p.shape[1] == 3:
                pf4 = pf4.to(device=dev, dtype=torch.float32).contiguous()
                np4 = np4.to(device=dev, dtype=torch.float32).contiguous()
                if np4.numel() != 4 * pf4.numel() or nrows[3] != 4 * (pf4.numel() // 3):
                    raise _lib.MolKGNNError("p_focal_deg4 / nei_p_deg4 / nei_edge_attr_deg4 disagree on the number of degree-4 nodes")
                keep += [pf4, np4]
            else:
                pf4 = np4 = None
            self._ref = (nea, nrows, pf4, np4, keep)
        self._pending = True
        return self

    def finish(self):
        if self.__dict__.get("_pending"):
            edge_index, p, edge_attr = self._keep
            self._pending = False
            if self._ref is not None:
                nea, nrows, pf4, np4, _ = self._ref
                check(_lib.lib().molkgnn_bucket_build_finish_ref(C.byref(self.c), ptr(edge_index), ptr(p), p.shape[1], nea, nrows,
                                                                 edge_attr.shape[1], ptr(pf4), ptr(np4), self.scratch_ptr,
                                                                 stream_ptr()))
            else:
                check(_lib.lib().molkgnn_bucket_build_finish(C.byref(self.c), ptr(edge_index), ptr(p), p.shape[1],
                                                             ptr(edge_attr), edge_attr.shape[1], self.scratch_ptr, stream_ptr()))
            self.n = list(self.c.n)
            self.n_tiles = int(self.c.n_tiles)      # 0: no molecule tiling (bucket-order kernels are used)
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
