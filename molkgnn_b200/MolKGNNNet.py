"""Drop-in for the reference's ``models/MolKGNN/MolKGNNNet.py`` (lines 10-149): the caller of the conv stack
(SURVEY 8(f) N1).  Same constructor, attribute / state-dict names and forward protocol:

    BatchNorm1d(x) -> MolGCN -> lin2(dropout(swish(lin1(h)))) -> global_add_pool        (MolKGNNNet.py:115-146)

The conv stack is the native one (one fused launch forward, csrc/stack_fwd_fused.cu); the head's two small dense layers are
plain library GEMMs (torch / cuBLAS), the pooling is the deterministic segmented sum of the C-ABI
(``molkgnn_segment_sum``: nodes of a graph are contiguous, added in node order -- torch's ``index_add_`` is atomic).
Like the reference, ``edge_batch_norm`` is evaluated (its running statistics advance) but its output is dead for the conv,
which reads the raw precomputed bond rows (kernels.py:679); ``graph_embedding_linear`` is never used (MolKGNNNet.py:20-25).
"""
from __future__ import annotations

import torch
from torch.nn import Linear, BatchNorm1d, Dropout

from . import _lib
from ._lib import ptr, stream_ptr, check
from .KernelLayer import MolGCN

DEG_KEYS = ("p_focal", "nei_p", "nei_edge_attr", "selected_index", "nei_index")


def swish(x):
    return x * x.sigmoid()                          # torch_geometric.nn.acts.swish


class _AddPool(torch.autograd.Function):
    """global_add_pool for a collated batch whose graphs are contiguous node ranges."""

    @staticmethod
    def forward(ctx, z, batch, ptr_):
        z = z.contiguous()
        B = ptr_.numel() - 1
        out = torch.empty(B, z.shape[1], dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            check(_lib.lib().molkgnn_segment_sum(ptr(z), z.shape[1], z.stride(0), ptr(ptr_), B, ptr(out), stream_ptr()))
        ctx.save_for_backward(batch)
        return out

    @staticmethod
    def backward(ctx, g):
        (batch,) = ctx.saved_tensors
        return g.index_select(0, batch), None, None


def global_add_pool(z, batch, ptr_=None):
    """PyG ``global_add_pool``.  ``ptr_`` ([B+1] graph boundaries) is derived from ``batch`` when the caller has none; a batch
    vector that is not sorted (graphs not contiguous) falls outside the collated-batch contract and raises."""
    if not z.is_cuda:
        raise _lib.MolKGNNError("molkgnn_b200.global_add_pool: CUDA tensors only (there is no CPU fallback)")
    if ptr_ is None:
        B = int(batch[-1]) + 1 if batch.numel() else 0
        if bool((batch[1:] < batch[:-1]).any()):
            raise _lib.MolKGNNError("global_add_pool: `batch` must be sorted (graphs contiguous, as PyG collation produces)")
        ptr_ = torch.zeros(B + 1, dtype=torch.int64, device=z.device)
        ptr_[1:] = torch.cumsum(torch.bincount(batch, minlength=B), 0)
    return _AddPool.apply(z.float(), batch, ptr_.to(torch.int64))


class MolKGNNNet(torch.nn.Module):
    def __init__(self, num_layers=1, num_kernel1_1hop=0, num_kernel2_1hop=0, num_kernel3_1hop=0, num_kernel4_1hop=0,
                 num_kernel1_Nhop=0, num_kernel2_Nhop=0, num_kernel3_Nhop=0, num_kernel4_Nhop=0,
                 predefined_kernelsets=True, x_dim=5, p_dim=3, edge_attr_dim=1, drop_ratio=0.25, graph_embedding_dim=5):
        super(MolKGNNNet, self).__init__()
        self.num_layers = num_layers
        self.D = p_dim
        K = num_kernel1_Nhop + num_kernel2_Nhop + num_kernel3_Nhop + num_kernel4_Nhop
        # same construction order as the reference (MolKGNNNet.py:20-56): parameter init consumes the RNG in this order
        self.graph_embedding_linear = Linear(K, graph_embedding_dim)
        self.node_batch_norm = BatchNorm1d(x_dim)
        self.edge_batch_norm = BatchNorm1d(edge_attr_dim)
        self.graph_embedding_lin1 = Linear(K, graph_embedding_dim)
        self.graph_embedding_lin2 = Linear(graph_embedding_dim, graph_embedding_dim)
        self.dropout = Dropout(drop_ratio)
        self.act = swish
        if self.num_layers < 1:
            raise ValueError("GNN_graphpred: Number of GNN layers must be greater than 0.")
        self.gnn = MolGCN(num_layers=num_layers, num_kernel1_1hop=num_kernel1_1hop, num_kernel2_1hop=num_kernel2_1hop,
                          num_kernel3_1hop=num_kernel3_1hop, num_kernel4_1hop=num_kernel4_1hop,
                          num_kernel1_Nhop=num_kernel1_Nhop, num_kernel2_Nhop=num_kernel2_Nhop,
                          num_kernel3_Nhop=num_kernel3_Nhop, num_kernel4_Nhop=num_kernel4_Nhop, x_dim=x_dim, p_dim=p_dim,
                          edge_attr_dim=edge_attr_dim)
        self.pool = global_add_pool

    def save_kernellayer(self, path, time_stamp):
        print(f'{self.D}D, there are {len(self.gnn.layers)} layers')
        self.gnn.save_kernellayer(path, time_stamp)

    def forward(self, *argv, save_score=False):
        if len(argv) == 1:
            data = argv[0]
        elif len(argv) == 33:
            # the reference unpacks 33 positional tensors here and then reads `data.x` (unbound: MolKGNNNet.py:115) -- that
            # branch cannot run there; it is rejected the same way it fails
            raise NameError("name 'data' is not defined (the reference's 33-argument branch never binds it, MolKGNNNet.py:73-115)")
        else:
            raise ValueError("unmatched number of arguments.")
        x = self.node_batch_norm(data.x)
        edge_attr = self.edge_batch_norm(data.edge_attr)     # dead for the conv, like the reference (statistics advance)
        kw = {}
        if all(hasattr(data, f"nei_edge_attr_deg{d}") for d in range(1, 5)):
            kw = {f"{k}_deg{d}": getattr(data, f"{k}_deg{d}", None) for d in range(1, 5) for k in DEG_KEYS}
            bond_rows = edge_attr
        else:
            # batches of molkgnn_b200.store (no precomputed per-degree tensors): the raw bond rows are gathered from the RAW
            # edge_attr, i.e. exactly what the reference's pre-transform would have stored (wrapper.py:578-593)
            bond_rows = data.edge_attr
        node_representation = self.gnn(x=x, edge_index=data.edge_index, edge_attr=bond_rows, p=data.p, save_score=save_score,
                                       **kw)
        z = self.graph_embedding_lin2(self.dropout(self.act(self.graph_embedding_lin1(node_representation))))
        return self.pool(z, data.batch, getattr(data, "ptr", None))
