import torch


def degree(index, num_nodes=None, dtype=None):
    """torch_geometric.utils.degree: out = zeros(N, dtype); out.scatter_add_(0, index, ones)."""
    n = int(num_nodes) if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)
