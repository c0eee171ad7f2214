import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(num_nodes) if num_nodes is not None else int(index.max()) + 1
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype, device=index.device))
