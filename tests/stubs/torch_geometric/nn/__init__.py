import torch
from . import acts  # noqa


class MessagePassing(torch.nn.Module):
    """aggr='add', flow source_to_target: out[i] = sum_{e: edge_index[1,e]==i} message(x[edge_index[0,e]])."""

    def __init__(self, aggr='add'):
        super().__init__()
        assert aggr == 'add'

    def propagate(self, edge_index, **kw):
        (name, val), = kw.items()
        msg = self.message(val.index_select(0, edge_index[0]))
        out = torch.zeros_like(val)
        return out.index_add(0, edge_index[1], msg)


def global_add_pool(x, batch):
    n = int(batch.max()) + 1 if batch.numel() else 0
    return torch.zeros(n, x.shape[1], dtype=x.dtype, device=x.device).index_add(0, batch, x)
