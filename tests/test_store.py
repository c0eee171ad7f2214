"""Packed molecule store, GPU batcher and the reader of the reference's on-disk formats (N3)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import molkgnn_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))


def _mols(n, seed):
    from molkgnn_b200 import synth
    return synth.make_molecules(n, seed=seed)


def _store(mols, y=None):
    from molkgnn_b200.store import MoleculeStore
    return MoleculeStore.from_molecules(mols, y=y)


def _check_same(st, mols):
    n = np.cumsum([0] + [m.num_nodes for m in mols])
    assert np.array_equal(st.node_ptr_host, n)
    assert np.array_equal(st.x.cpu().numpy(), np.concatenate([m.x for m in mols]))
    assert np.array_equal(st.edge_index.cpu().numpy(), np.concatenate([m.edge_index for m in mols], axis=1))


def test_reader_of_the_reference_dataset_format_pyg1_style(tmp_path):
    """kgnn-{AID}-3D.pt = torch.save((data, slices)) (wrapper.py:449-450), attributes in Data.__dict__ (PyG 1.x / the stub)."""
    from torch_geometric.data import Data            # tests/stubs: attribute bag
    from molkgnn_b200.store import MoleculeStore
    mols = _mols(7, 4)
    y = np.arange(7, dtype=np.float32) % 2
    st = _store(mols, y=y)
    path = str(tmp_path / "kgnn-435008-3D.pt")
    st.save_reference_pt(path, data_cls=Data)
    back = MoleculeStore.from_reference_pt(path)
    _check_same(back, mols)
    assert np.array_equal(back.y.numpy().reshape(-1), y)


def test_reader_of_the_reference_dataset_format_pyg2_style(tmp_path):
    """PyG >= 2.0 (the reference pins 2.0.4): Data keeps its attributes in _store (GlobalStorage._mapping).  The pickle is
    written with stand-in classes of those module paths and read back WITHOUT any torch_geometric module importable."""
    from molkgnn_b200.store import MoleculeStore
    mols = _mols(5, 9)
    st = _store(mols)
    mod_d, mod_s = types.ModuleType("torch_geometric.data.data"), types.ModuleType("torch_geometric.data.storage")

    class GlobalStorage(object):
        def __init__(self, mapping):
            self._mapping = mapping

    class Data(object):
        def __init__(self, **kw):
            self._store = GlobalStorage(dict(kw))

    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mod_d.Data, mod_s.GlobalStorage = Data, GlobalStorage
    saved = {k: sys.modules.get(k) for k in ("torch_geometric.data.data", "torch_geometric.data.storage")}
    sys.modules["torch_geometric.data.data"], sys.modules["torch_geometric.data.storage"] = mod_d, mod_s
    path = str(tmp_path / "kgnn-1798-3D.pt")
    try:
        st.save_reference_pt(path, data_cls=Data)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    back = MoleculeStore.from_reference_pt(path)
    _check_same(back, mols)


def test_split_files_of_the_reference():
    from molkgnn_b200.store import load_split
    path = "/root/reference/data_split/1798_seed2.pt"
    if not os.path.exists(path):
        pytest.skip("/root/reference is not present on this box")
    sp = load_split(path)
    assert set(sp) == {"train", "valid", "test"} and sp["train"].dtype == torch.int64
    assert len(sp["train"]) == 49466 and len(sp["valid"]) == 6183 and len(sp["test"]) == 6182
    allids = torch.cat(list(sp.values()))
    assert len(torch.unique(allids)) == len(allids)


def test_cpu_store_refuses_to_collate():
    import molkgnn_b200 as mk
    st = _store(_mols(3, 1))
    with pytest.raises(mk.MolKGNNError):
        st.collate([0, 1])


@pytest.mark.gpu
def test_collate_bit_exact_with_pyg_semantics():
    mols = _mols(200, 11)
    y = np.random.default_rng(0).standard_normal((200, 1)).astype(np.float32)
    st = _store(mols, y=y).to("cuda")
    rng = np.random.default_rng(5)
    for ids in (np.arange(200), rng.permutation(200)[:37], rng.integers(0, 200, 300), np.array([199]), np.array([5, 5, 5])):
        b = st.collate(ids)
        ref = orc.collate_pyg([mols[i] for i in ids])
        for k in ("x", "p", "edge_attr", "edge_index", "batch", "ptr"):
            got = b[k].cpu().numpy()
            assert got.shape == ref[k].shape and got.dtype == ref[k].dtype, (k, got.shape, ref[k].shape)
            assert np.array_equal(got, ref[k]), k
        assert np.array_equal(b["y"].cpu().numpy(), y[ids])


@pytest.mark.gpu
def test_loader_feeds_the_conv_stack():
    """One epoch of StoreLoader batches through MolGCN == the same molecules collated on the host."""
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth
    from molkgnn_b200.store import StoreLoader
    mols = _mols(50, 21)
    st = _store(mols).to("cuda")
    torch.manual_seed(0)
    net = mk.MolGCN(2, 3, 4, 5, 6, 2, 3, 4, 5, x_dim=28, p_dim=3, edge_attr_dim=7).to("cuda")
    loader = StoreLoader(st, batch_size=16, shuffle=False)
    seen = 0
    with torch.no_grad():
        for i, b in enumerate(loader):
            h = net(x=b["x"], edge_index=b["edge_index"], edge_attr=b["edge_attr"], p=b["p"], save_score=False)
            ref = synth.collate(mols[16 * i:16 * (i + 1)])
            t = {k: torch.from_numpy(ref[k]).cuda() for k in ("x", "p", "edge_index", "edge_attr")}
            h2 = net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False)
            assert torch.equal(h, h2)
            seen += int(b["ptr"].numel()) - 1
    assert seen == 50 and len(loader) == 4
    w = np.where(np.arange(50) % 10 == 0, 10.0, 1.0)          # over-sampling with replacement (data.py:150-167)
    assert sum(int(b["ptr"].numel()) - 1 for b in StoreLoader(st, batch_size=16, weights=w, seed=3)) == 50
