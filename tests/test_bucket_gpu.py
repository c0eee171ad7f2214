"""GPU: the CUDA degree-bucket pass against the reference transform's golden output and the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import molkgnn_oracle as orc
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
KEYS = ["selected_index", "nei_index", "p_focal", "nei_p", "nei_edge_attr"]


def _plan(ei, p, ea, n):
    from molkgnn_b200 import BucketPlan
    dev = torch.device("cuda")
    return BucketPlan.from_edge_index(torch.from_numpy(ei).to(dev), torch.from_numpy(p).to(dev),
                                      torch.from_numpy(ea).to(dev), n)


@pytest.mark.parametrize("name", ["bucket_a", "bucket_b"])
def test_bucket_golden_bit_exact(name):
    g = load_golden(name)
    plan = _plan(g["edge_index"], g["p"], g["edge_attr"], int(g["num_nodes"]))
    p, ea = torch.from_numpy(g["p"]).cuda(), torch.from_numpy(g["edge_attr"]).cuda()
    for d in range(1, 5):
        ref_n = g[f"selected_index_deg{d}"].shape[0]
        assert plan.n[d - 1] == ref_n
        if ref_n == 0:
            continue
        out = plan.export(d, p, ea)
        for k in KEYS:
            ref = g[f"{k}_deg{d}"]
            got = out[k].cpu().numpy()
            assert got.dtype == ref.dtype, (k, got.dtype, ref.dtype)
            assert got.shape == ref.shape, (k, d, got.shape, ref.shape)
            assert np.array_equal(got, ref), (k, d)


@pytest.mark.parametrize("n_mol,seed", [(1, 0), (64, 1), (1024, 2)])
def test_bucket_vs_oracle(n_mol, seed):
    from molkgnn_b200 import synth
    b = synth.make_batch(n_mol, seed=seed)
    N = b["x"].shape[0]
    ref = orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"])
    plan = _plan(b["edge_index"], b["p"], b["edge_attr"], N)
    assert np.array_equal(plan.deg.cpu().numpy(), ref["deg"])
    p, ea = torch.from_numpy(b["p"]).cuda(), torch.from_numpy(b["edge_attr"]).cuda()
    for d in range(1, 5):
        assert plan.n[d - 1] == ref[d]["selected_index"].shape[0]
        if plan.n[d - 1] == 0:
            continue
        out = plan.export(d, p, ea)
        for k in KEYS:
            assert np.array_equal(out[k].cpu().numpy(), ref[d][k]), (k, d)
    # in-lists: sources of in-edges in edge order + position inside the source's neighbour list
    ei = b["edge_index"]
    order = np.argsort(ei[1], kind="stable")
    in_src = plan.in_src.cpu().numpy()
    in_cnt = plan.in_cnt.cpu().numpy()
    assert np.array_equal(in_cnt, np.bincount(ei[1], minlength=N))
    ptr = np.concatenate([[0], np.cumsum(in_cnt)])
    in_j = plan.in_j.cpu().numpy()
    oorder = np.argsort(ei[0], kind="stable")
    optr = np.concatenate([[0], np.cumsum(ref["deg"])])
    sample = set(np.random.default_rng(0).integers(0, N, size=min(N, 200)).tolist()) | set(range(min(N, 40)))
    for v in sample:
        eids = order[ptr[v]:ptr[v + 1]]
        assert np.array_equal(in_src[v, :in_cnt[v]], ei[0][eids])
        for t, e in enumerate(eids):       # in_j = rank of the edge inside its source's (edge ordered) out-list
            u = ei[0][e]
            assert oorder[optr[u] + in_j[v, t]] == e, (v, t)


def test_shuffled_edge_order():
    """edge_index that is not grouped by source: neighbours must still come out in edge order (wrapper.py:567-572)."""
    from molkgnn_b200 import synth
    b = synth.make_batch(32, seed=5)
    rng = np.random.default_rng(1)
    E = b["edge_index"].shape[1]
    # permute BONDS (pairs of rows) so that the 2*(eid//2) bond-row convention still holds
    perm = rng.permutation(E // 2)
    idx = np.stack([2 * perm, 2 * perm + 1], 1).reshape(-1)
    ei, ea = b["edge_index"][:, idx], b["edge_attr"][idx]
    N = b["x"].shape[0]
    ref = orc.bucket_pass(ei, N, b["p"], ea)
    plan = _plan(np.ascontiguousarray(ei), b["p"], np.ascontiguousarray(ea), N)
    p, eat = torch.from_numpy(b["p"]).cuda(), torch.from_numpy(np.ascontiguousarray(ea)).cuda()
    for d in range(1, 5):
        if plan.n[d - 1] == 0:
            continue
        out = plan.export(d, p, eat)
        for k in KEYS:
            assert np.array_equal(out[k].cpu().numpy(), ref[d][k]), (k, d)


def test_transform_drop_in():
    from molkgnn_b200 import ToXAndPAndEdgeAttrForDeg, synth

    class Bag(object):
        pass
    g = load_golden("bucket_b")
    d = Bag()
    d.x = torch.zeros(int(g["num_nodes"]), 28)
    d.p, d.edge_index, d.edge_attr = (torch.from_numpy(g[k]) for k in ["p", "edge_index", "edge_attr"])
    d = ToXAndPAndEdgeAttrForDeg()(d)
    for deg in range(1, 5):
        for k in KEYS:
            got = getattr(d, f"{k}_deg{deg}")
            assert not got.is_cuda
            if g[f"{k}_deg{deg}"].size:
                assert np.array_equal(got.numpy(), g[f"{k}_deg{deg}"])


def test_rejects_bad_degree():
    from molkgnn_b200 import BucketPlan
    from molkgnn_b200._lib import MolKGNNError
    # star with 5 leaves: centre has degree 5
    src = [0, 1, 0, 2, 0, 3, 0, 4, 0, 5]
    dst = [1, 0, 2, 0, 3, 0, 4, 0, 5, 0]
    ei = torch.tensor([src, dst], dtype=torch.int64).cuda()
    with pytest.raises(MolKGNNError):
        BucketPlan.from_edge_index(ei, torch.zeros(6, 3).cuda(), torch.ones(10, 7).cuda(), 6)
    # isolated node (degree 0)
    ei = torch.tensor([[0, 1], [1, 0]], dtype=torch.int64).cuda()
    with pytest.raises(MolKGNNError):
        BucketPlan.from_edge_index(ei, torch.zeros(3, 3).cuda(), torch.ones(2, 7).cuda(), 3)


@pytest.mark.parametrize("n_mol,seed", [(1, 0), (7, 3), (300, 1), (4096, 2)])
def test_molecule_tiles(n_mol, seed):
    """Tiles are consecutive node ranges of whole molecules (no edge crosses a boundary), at most 128 nodes each."""
    from molkgnn_b200 import synth
    b = synth.make_batch(n_mol, seed=seed)
    N = b["x"].shape[0]
    plan = _plan(b["edge_index"], b["p"], b["edge_attr"], N)
    T = plan.n_tiles
    assert T > 0
    ts = plan.tile_start.cpu().numpy()[:T + 1]
    assert ts[0] == 0 and ts[T] == N
    assert np.all(np.diff(ts) >= 0) and np.diff(ts).max() <= 128
    assert plan.c.tile_max_nodes == np.diff(ts).max()
    tile_of = np.searchsorted(ts[1:], np.arange(N), side="right")
    ei = b["edge_index"]
    assert np.array_equal(tile_of[ei[0]], tile_of[ei[1]])
    deg = plan.deg.cpu().numpy()
    for d in range(1, 5):
        cnt = np.bincount(tile_of[deg == d], minlength=T)
        assert plan.c.tile_max_deg[d - 1] == cnt.max()
    # greedy packing: a tile ends only where the next molecule would not fit any more (batch['batch'] = molecule of a node)
    mol = b["batch"]
    first = np.flatnonzero(np.diff(np.concatenate([[-1], mol])))           # first node of every molecule
    sizes = np.diff(np.concatenate([first, [N]]))
    for t in range(T - 1):
        m = mol[ts[t + 1]]                                                 # molecule that opens the next tile
        assert ts[t + 1] == first[m]
        assert ts[t + 1] - ts[t] + sizes[m] > 128
    if n_mol >= 300:
        assert np.diff(ts).mean() > 105


@pytest.mark.parametrize("n_mol,seed", [(1, 0), (7, 3), (300, 1), (4096, 2), (4096, 6000)])
def test_tile_schedule_is_a_balanced_permutation(n_mol, seed):
    """plan.tile_order (k_tile_order): every tile exactly once; CTA b of the G-CTA tile-major kernels walks a contiguous
    slice of q or q + 1 entries (csrc/tile.cuh TileWalk); the per-CTA node totals are much closer than round robin's; the
    schedule is deterministic.  Seed 6000 is the 8-GPU shard whose 892 tiles overflow 6 rounds of 148 SMs by four."""
    from molkgnn_b200 import synth, _lib
    b = synth.make_batch(n_mol, seed=seed)
    N = b["x"].shape[0]
    sms = _lib.lib().molkgnn_num_sms()
    plan = _plan(b["edge_index"], b["p"], b["edge_attr"], N)          # default mode: only where a few CTAs walk one tile more
    rem = plan.n_tiles % min(plan.n_tiles, sms)
    assert plan.c.tile_grid == (min(plan.n_tiles, sms) if 0 < rem <= min(plan.n_tiles, sms) // 2 else 0)
    old = _lib.lib().molkgnn_set_tile_order(2)                         # always
    try:
        plan = _plan(b["edge_index"], b["p"], b["edge_attr"], N)
        plan2 = _plan(b["edge_index"], b["p"], b["edge_attr"], N)
    finally:
        _lib.lib().molkgnn_set_tile_order(old)
    T, G = plan.n_tiles, plan.c.tile_grid
    assert G == min(T, sms)
    order = plan.tile_order.cpu().numpy()[:T]
    assert np.array_equal(np.sort(order), np.arange(T))
    assert np.array_equal(plan2.tile_order.cpu().numpy()[:T], order)
    nn = np.diff(plan.tile_start.cpu().numpy()[:T + 1])
    from tests.helpers import tile_schedule_model
    assert np.array_equal(order, tile_schedule_model(nn, G))           # the kernel against its plain-numpy model
    q, r = divmod(T, G)
    load, load_rr = np.zeros(G), np.zeros(G)
    pos = 0
    for c in range(G):
        cnt = q + (1 if c < r else 0)
        load[c] = nn[order[pos:pos + cnt]].sum()
        pos += cnt
    for t in range(T):
        load_rr[t % G] += nn[t]
    assert pos == T
    if T >= 4 * G:
        assert load.max() / load.mean() < 1.03
        assert load.max() <= load_rr.max()
