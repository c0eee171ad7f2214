"""Packed molecule store + GPU-side batcher and reader of the reference's processed datasets (N3).

Reference path that this replaces (host side, per step):
  * ``data.py:136-229``   torch_geometric ``DataLoader``s over an ``InMemoryDataset``: worker processes slice every one of the
                          30 attributes of each molecule out of the collated store and PyG's ``Batch.from_data_list`` glues them
                          back together (``__cat_dim__`` / ``__inc__`` semantics), then Lightning copies the batch to the GPU;
  * ``wrapper.py:392, 449-450``   the processed dataset ``kgnn-{AID}-3D.pt`` = ``torch.save((data, slices))``;
  * ``wrapper.py:510``    ``data_split/shrink_{AID}_seed2.pt`` = dict ``{'train','valid','test'}`` of molecule-id lists.

Here the whole dataset lives in HBM once (180 GB: the largest set, AID 435008 with 218 k molecules, is < 1 GB), packed back
to back with CSR pointers, and a batch is assembled by ONE C-ABI call (two kernels, csrc/collate.cu) from a list of molecule
ids -- no worker processes, no host->device copy of features per step.  The 20 per-degree attributes the reference stores
(``p_focal_deg*`` ...) are NOT stored: the GPU bucket pass rebuilds them per batch, bit-exactly (tests/test_bucket_gpu.py).
"""
from __future__ import annotations

import ctypes as C
import io
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream_ptr, check

FLOAT_KEYS = ("x", "p", "edge_attr")


class _Bag(object):
    """Stand-in for any class of a pickled PyG object graph (torch_geometric is not needed to READ a processed dataset)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state
        if isinstance(state, dict):
            self.__dict__.update(state)


class _TolerantUnpickler(pickle.Unpickler):
    """Resolves torch / numpy / builtins normally and maps every ``torch_geometric.*`` class to an attribute bag."""

    def find_class(self, module, name):
        if module.split(".")[0] == "torch_geometric":
            return type(name, (_Bag,), {"__module__": module})
        return super().find_class(module, name)


class _TolerantPickle(object):
    Unpickler = _TolerantUnpickler
    __name__ = "pickle"

    @staticmethod
    def load(f, **kw):
        return _TolerantUnpickler(f, **kw).load()


def _find_mapping(obj, depth=0):
    """Finds the attribute mapping of a (stubbed) PyG ``Data`` object: PyG >= 2.0 keeps it in ``_store._mapping``, PyG 1.x in
    ``__dict__`` -- search for the first dict that holds an ``edge_index`` tensor."""
    if isinstance(obj, dict):
        if "edge_index" in obj and torch.is_tensor(obj["edge_index"]):
            return obj
        vals = obj.values()
    elif hasattr(obj, "__dict__"):
        vals = obj.__dict__.values()
    else:
        return None
    if depth > 4:
        return None
    for v in vals:
        if isinstance(v, (dict, _Bag)) or hasattr(v, "__dict__"):
            r = _find_mapping(v, depth + 1)
            if r is not None:
                return r
    return None


def load_split(path):
    """``data_split/*.pt`` of the reference (wrapper.py:510, utils/data_split.py): dict of id lists -> dict of int64 tensors."""
    d = torch.load(path, weights_only=False)
    return {k: torch.as_tensor(v, dtype=torch.int64) for k, v in d.items()}


class MoleculeStore(object):
    """All molecules of a dataset, packed: ``x [sumN,F]``, ``p [sumN,P]``, ``edge_attr [sumE,Fe]``, ``edge_index [2,sumE]``
    (molecule-LOCAL ids), ``y [M,Y]`` (optional), ``node_ptr`` / ``edge_ptr [M+1]`` (int64; also kept on the host)."""

    def __init__(self, x, p, edge_index, edge_attr, node_ptr, edge_ptr, y=None):
        self.x, self.p, self.edge_attr = x.float().contiguous(), p.float().contiguous(), edge_attr.float().contiguous()
        self.edge_index = edge_index.long().contiguous()
        self.node_ptr, self.edge_ptr = node_ptr.long().contiguous(), edge_ptr.long().contiguous()
        self.y = None if y is None else y.float().reshape(len(node_ptr) - 1, -1).contiguous()
        self.node_ptr_host = self.node_ptr.cpu().numpy()
        self.edge_ptr_host = self.edge_ptr.cpu().numpy()
        self.num_molecules = len(self.node_ptr_host) - 1
        self._err = None
        if self.x.shape[0] != self.node_ptr_host[-1] or self.edge_index.shape[1] != self.edge_ptr_host[-1]:
            raise _lib.MolKGNNError("MoleculeStore: pointer arrays do not match the packed tensors")

    def __len__(self):
        return self.num_molecules

    @property
    def device(self):
        return self.x.device

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)  # noqa: E731
        return MoleculeStore(mv(self.x), mv(self.p), mv(self.edge_index), mv(self.edge_attr), mv(self.node_ptr),
                             mv(self.edge_ptr), mv(self.y))

    # ---- constructors ---------------------------------------------------------------------------------------------
    @classmethod
    def from_molecules(cls, mols, y=None):
        """``mols``: objects with x [n,F], p [n,P], edge_index [2,e] (local ids), edge_attr [e,Fe] (numpy or torch)."""
        t = lambda a: torch.as_tensor(np.asarray(a))  # noqa: E731
        n = np.array([np.asarray(m.x).shape[0] for m in mols], dtype=np.int64)
        e = np.array([np.asarray(m.edge_index).shape[1] for m in mols], dtype=np.int64)
        return cls(torch.cat([t(m.x) for m in mols]), torch.cat([t(m.p) for m in mols]),
                   torch.cat([t(m.edge_index) for m in mols], dim=1), torch.cat([t(m.edge_attr) for m in mols]),
                   torch.from_numpy(np.concatenate([[0], np.cumsum(n)])), torch.from_numpy(np.concatenate([[0], np.cumsum(e)])),
                   None if y is None else torch.as_tensor(np.asarray(y)))

    @classmethod
    def from_data_slices(cls, data, slices):
        """From the reference's in-memory representation (PyG ``InMemoryDataset``: ``data`` holds every attribute
        concatenated along its cat dim WITHOUT index increments, ``slices[key]`` the boundaries): wrapper.py:449."""
        g = data if isinstance(data, dict) else _find_mapping(data)
        if g is None:
            raise _lib.MolKGNNError("MoleculeStore: cannot find the attribute mapping of the dataset's Data object")
        for k in ("x", "p", "edge_index", "edge_attr"):
            if k not in g or k not in slices:
                raise _lib.MolKGNNError(f"MoleculeStore: the dataset has no attribute '{k}'")
        node_ptr, edge_ptr = torch.as_tensor(slices["x"]).long(), torch.as_tensor(slices["edge_index"]).long()
        if not torch.equal(node_ptr, torch.as_tensor(slices["p"]).long()) or not torch.equal(
                edge_ptr, torch.as_tensor(slices["edge_attr"]).long()):
            raise _lib.MolKGNNError("MoleculeStore: inconsistent slices")
        y = g.get("y", None)
        return cls(g["x"], g["p"], g["edge_index"], g["edge_attr"], node_ptr, edge_ptr,
                   None if y is None or not torch.is_tensor(y) else y)

    @classmethod
    def from_reference_pt(cls, path):
        """Reads ``kgnn-{AID}-3D.pt`` as written by the reference (``torch.save((data, slices))``, wrapper.py:449-450) without
        torch_geometric: PyG classes in the pickle are mapped to attribute bags."""
        data, slices = torch.load(path, pickle_module=_TolerantPickle, weights_only=False)
        return cls.from_data_slices(data, slices)

    def save_reference_pt(self, path, data_cls=None):
        """Writes the same (data, slices) tuple layout (plain dict as ``data`` unless a PyG-like ``data_cls`` is given)."""
        g = dict(x=self.x.cpu(), p=self.p.cpu(), edge_index=self.edge_index.cpu(), edge_attr=self.edge_attr.cpu())
        if self.y is not None:
            g["y"] = self.y.cpu()
        sl = dict(x=self.node_ptr.cpu(), p=self.node_ptr.cpu(), edge_index=self.edge_ptr.cpu(), edge_attr=self.edge_ptr.cpu())
        if self.y is not None:
            sl["y"] = torch.arange(self.num_molecules + 1)
        torch.save((g if data_cls is None else data_cls(**g), sl), path)

    # ---- batch assembly -------------------------------------------------------------------------------------------------
    def collate(self, ids):
        """Batch of the molecules ``ids`` (sequence / int64 tensor; any order, repeats allowed) -> dict with the collated
        ``x, p, edge_index, edge_attr, batch, ptr`` (+ ``y``) on the store's device.  PyG ``Batch.from_data_list`` semantics."""
        if not self.x.is_cuda:
            raise _lib.MolKGNNError("MoleculeStore.collate: the store must live on a CUDA device (store.to('cuda')); "
                                    "there is no CPU fallback")
        if torch.is_tensor(ids):
            ids_host = (ids if not ids.is_cuda else ids.cpu()).numpy().astype(np.int64, copy=False).reshape(-1)
        else:
            ids_host = np.asarray(ids, dtype=np.int64).reshape(-1)
        M = int(ids_host.shape[0])
        if M == 0:
            raise _lib.MolKGNNError("MoleculeStore.collate: empty batch")
        if ids_host.min() < 0 or ids_host.max() >= self.num_molecules:
            raise _lib.MolKGNNError("MoleculeStore.collate: molecule id out of range")
        # batch sizes from the host copies of the pointers: no device round trip
        Nb = int((self.node_ptr_host[ids_host + 1] - self.node_ptr_host[ids_host]).sum())
        Eb = int((self.edge_ptr_host[ids_host + 1] - self.edge_ptr_host[ids_host]).sum())
        dev = self.device
        with torch.cuda.device(dev):
            if torch.is_tensor(ids) and ids.is_cuda:
                ids_dev = ids.to(dev, torch.int64).contiguous()
            elif torch.is_tensor(ids) and ids.dtype == torch.int64 and ids.is_pinned() and ids.is_contiguous():
                ids_dev = ids.to(dev, non_blocking=True)          # the caller's pinned id buffer: the step's only H2D copy
            else:
                ids_dev = torch.from_numpy(ids_host).pin_memory().to(dev, non_blocking=True)
            F, P, Fe = self.x.shape[1], self.p.shape[1], self.edge_attr.shape[1]
            Y = 0 if self.y is None else self.y.shape[1]
            # ONE allocation for the whole batch, carved into the output tensors (16-byte aligned pieces)
            al = lambda n: (n + 15) // 16 * 16  # noqa: E731
            sizes = [("x", Nb * F * 4), ("p", Nb * P * 4), ("edge_attr", Eb * Fe * 4), ("edge_index", 2 * Eb * 8),
                     ("batch", Nb * 8), ("ptr", (M + 1) * 8), ("y", M * Y * 4)]
            offs, tot = {}, 0
            for k, nb in sizes:
                offs[k] = tot
                tot += al(nb)
            buf = torch.empty(tot, dtype=torch.uint8, device=dev)
            cut = lambda k, nb, dt, shape: buf[offs[k]:offs[k] + nb].view(dt).view(*shape)  # noqa: E731
            out = dict(x=cut("x", Nb * F * 4, torch.float32, (Nb, F)), p=cut("p", Nb * P * 4, torch.float32, (Nb, P)),
                       edge_attr=cut("edge_attr", Eb * Fe * 4, torch.float32, (Eb, Fe)),
                       edge_index=cut("edge_index", 2 * Eb * 8, torch.int64, (2, Eb)),
                       batch=cut("batch", Nb * 8, torch.int64, (Nb,)), ptr=cut("ptr", (M + 1) * 8, torch.int64, (M + 1,)))
            if Y:
                out["y"] = cut("y", M * Y * 4, torch.float32, (M, Y))
            scratch = torch.empty(2 * (M + 1), dtype=torch.int64, device=dev)
            if self._err is None:
                self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            check(_lib.lib().molkgnn_collate(ptr(ids_dev), M, self.num_molecules, ptr(self.node_ptr), ptr(self.edge_ptr),
                                             ptr(self.x), F, ptr(self.p), P, ptr(self.edge_attr), Fe, ptr(self.edge_index),
                                             self.edge_index.shape[1], ptr(self.y), Y, ptr(scratch), ptr(scratch[M + 1:]),
                                             ptr(out["x"]), ptr(out["p"]), ptr(out["edge_attr"]), ptr(out["edge_index"]), Eb,
                                             ptr(out["batch"]), ptr(out["ptr"]), ptr(out.get("y")), ptr(self._err), stream_ptr()))
            out["_keep"] = (ids_dev, scratch)
        return out


class StoreLoader(object):
    """Iterates collated GPU batches over a subset of the store -- the role of ``DataLoaderModule.train_dataloader`` /
    ``val_dataloader`` (data.py:136-216): ``shuffle`` = random permutation per epoch, ``weights`` = ``WeightedRandomSampler``
    with replacement (the reference's over-sampling of actives, data.py:150-167), else sequential."""

    def __init__(self, store, ids=None, batch_size=16, shuffle=False, weights=None, seed=0, drop_last=False):
        self.store = store
        self.ids = torch.arange(len(store)) if ids is None else torch.as_tensor(ids, dtype=torch.int64)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.weights = None if weights is None else torch.as_tensor(weights, dtype=torch.double)
        self.gen = torch.Generator()
        self.gen.manual_seed(seed)

    def __len__(self):
        n = len(self.ids)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.ids)
        if self.weights is not None:
            order = torch.multinomial(self.weights, n, replacement=True, generator=self.gen)
        elif self.shuffle:
            order = torch.randperm(n, generator=self.gen)
        else:
            order = torch.arange(n)
        for i in range(len(self)):
            yield self.store.collate(self.ids[order[i * self.batch_size:(i + 1) * self.batch_size]])
