"""Seeded synthetic 3D drug-like molecules + PyG-semantics batch collation.

Shape contract follows the reference featuriser ``mol2graph`` (/root/reference/wrapper.py:103-167):
``x[n,28] f32``, ``p[n,3] f32``, ``edge_index[2,2*bonds] i64`` with bond ``b`` on rows ``2b:(i,j)`` and
``2b+1:(j,i)`` (wrapper.py:152-156), ``edge_attr[2*bonds,7] f32`` duplicated for both directions.
The generator itself is the one specified in SURVEY.md 8(d): random tree grown atom by atom with
degree-weighted parent choice, 2..4 ring closures between atoms of degree <= 2, all degrees in 1..4.

Collation reproduces what PyG 2.0.x ``Batch.from_data_list`` does for the attributes the conv path
reads (``Data.__inc__``/``__cat_dim__``): keys containing ``index`` are offset by the cumulative node
count and concatenated along the last dim, everything else along dim 0.
"""
from __future__ import annotations

import numpy as np

X_DIM = 28
EDGE_DIM = 7
_PARENT_WEIGHT = np.array([6.0, 6.0, 1.0, 2.0, 0.0])  # by current degree 0..4


class Molecule(object):
    __slots__ = ("x", "p", "edge_index", "edge_attr")

    def __init__(self, x, p, edge_index, edge_attr):
        self.x, self.p, self.edge_index, self.edge_attr = x, p, edge_index, edge_attr

    @property
    def num_nodes(self):
        return self.x.shape[0]


def _unit(rng):
    v = rng.standard_normal(3)
    return v / max(np.linalg.norm(v), 1e-12)


def make_molecule(rng: np.random.Generator, min_atoms=18, max_atoms=32, dup_leaf_prob=0.5) -> Molecule:
    n = int(rng.integers(min_atoms, max_atoms + 1))
    deg = np.zeros(n, dtype=np.int64)
    parent = np.full(n, -1, dtype=np.int64)
    pos = np.zeros((n, 3), dtype=np.float64)
    bonds = []
    adj = [set() for _ in range(n)]
    for v in range(1, n):
        w = _PARENT_WEIGHT[deg[:v]]
        if w.sum() <= 0:  # cannot happen for n<=32 trees but keep the generator total
            w = (deg[:v] < 4).astype(np.float64)
        u = int(rng.choice(v, p=w / w.sum()))
        parent[v] = u
        bonds.append((u, v))
        adj[u].add(v)
        adj[v].add(u)
        deg[u] += 1
        deg[v] += 1
        pos[v] = pos[u] + 1.5 * _unit(rng)
    n_ring = int(rng.integers(2, 5))
    made, attempts = 0, 0
    while made < n_ring and attempts < 200:
        attempts += 1
        a, b = (int(t) for t in rng.integers(0, n, size=2))
        if a == b or deg[a] > 2 or deg[b] > 2 or b in adj[a]:
            continue
        bonds.append((a, b))
        adj[a].add(b)
        adj[b].add(a)
        deg[a] += 1
        deg[b] += 1
        made += 1
    x = rng.standard_normal((n, X_DIM)).astype(np.float32)
    # leaves hanging off the same parent share one feature row with prob. dup_leaf_prob: exercises the
    # tied-permutation and chirality-gate paths already at layer 0 (kernels.py:310-317)
    leaves_of = {}
    for v in range(1, n):
        if deg[v] == 1:
            leaves_of.setdefault(int(parent[v]), []).append(v)
    for u, ls in leaves_of.items():
        if len(ls) >= 2 and rng.random() < dup_leaf_prob:
            for v in ls[1:]:
                x[v] = x[ls[0]]
    nb = len(bonds)
    ea = np.zeros((nb, EDGE_DIM), dtype=np.float32)
    ea[np.arange(nb), rng.integers(0, 4, size=nb)] = 1.0
    ea[:, 4:] = (rng.random((nb, 3)) < 0.5).astype(np.float32)
    edge_index = np.empty((2, 2 * nb), dtype=np.int64)
    b = np.asarray(bonds, dtype=np.int64)
    edge_index[0, 0::2], edge_index[1, 0::2] = b[:, 0], b[:, 1]
    edge_index[0, 1::2], edge_index[1, 1::2] = b[:, 1], b[:, 0]
    edge_attr = np.repeat(ea, 2, axis=0)
    return Molecule(x, pos.astype(np.float32), edge_index, edge_attr)


def make_molecules(num, seed=0, **kw):
    rng = np.random.Generator(np.random.PCG64(seed))
    return [make_molecule(rng, **kw) for _ in range(num)]


def collate(mols):
    """PyG-semantics collation of raw graphs -> dict of numpy arrays (x, p, edge_index, edge_attr, batch, ptr)."""
    n_nodes = np.array([m.num_nodes for m in mols], dtype=np.int64)
    ptr = np.concatenate([[0], np.cumsum(n_nodes)])
    x = np.concatenate([m.x for m in mols], axis=0)
    p = np.concatenate([m.p for m in mols], axis=0)
    edge_index = np.concatenate([m.edge_index + off for m, off in zip(mols, ptr[:-1])], axis=1)
    edge_attr = np.concatenate([m.edge_attr for m in mols], axis=0)
    batch = np.repeat(np.arange(len(mols), dtype=np.int64), n_nodes)
    return dict(x=x, p=p, edge_index=edge_index, edge_attr=edge_attr, batch=batch, ptr=ptr)


def make_batch(num_molecules, seed=0, pool=None, **kw):
    """Collated batch of ``num_molecules`` synthetic molecules.

    For very large batches (multi-GPU shards, the screening sweep) pure-Python graph growth is the
    bottleneck, so ``pool`` topologies are grown once and sampled with replacement; node features,
    bond attributes' flags and coordinates jitter are redrawn per instance so no two molecules are equal.
    """
    if pool is None or pool >= num_molecules:
        return collate(make_molecules(num_molecules, seed=seed, **kw))
    rng = np.random.Generator(np.random.PCG64(seed))
    base = [make_molecule(rng, **kw) for _ in range(pool)]
    pick = rng.integers(0, pool, size=num_molecules)
    out = []
    for i in pick:
        m = base[int(i)]
        # keep duplicate-leaf structure: redraw rows, then re-tie rows that were tied in the template
        _, first = np.unique(m.x, axis=0, return_inverse=True)
        fresh = rng.standard_normal((first.max() + 1, X_DIM)).astype(np.float32)
        x = fresh[first.reshape(-1)]
        p = m.p + (0.05 * rng.standard_normal(m.p.shape)).astype(np.float32)
        out.append(Molecule(x, p, m.edge_index, m.edge_attr))
    return collate(out)
