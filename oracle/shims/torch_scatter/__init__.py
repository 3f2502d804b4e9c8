"""Shim of torch_scatter (rusty1s/pytorch_scatter >= 2.0) -- only `scatter`, only what
reference models/pytorch_geometric/aggregators.py:10,14,18,22 calls: scatter(src, index, 0, None, dim_size, reduce=...).

Upstream behaviour restated (torch_scatter/scatter.py):
  scatter_sum : out = zeros(size); out.scatter_add_(dim, broadcast(index, src, dim), src)
  scatter_mean: out = scatter_sum(src); count = scatter_sum(ones(index.size())); count.clamp_(1);
                out.true_divide_(broadcast(count, out, dim))   (floating point)
  scatter_min / scatter_max: C++ kernel; `out` is filled with the reducer's init value, reduced, and entries no index
                pointed at are set to 0 when `out` was not supplied (csrc/cpu/scatter_cpu.cpp:
                `out.masked_fill_(arg_out == src.size(dim), 0)`).
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None, "shim: `out` is never passed by the reference"
    bidx = _broadcast(index, src, dim)
    size = list(src.size())
    size[dim] = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    if reduce in ("sum", "add"):
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, bidx, src)
    if reduce == "mean":
        res = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, bidx, src)
        ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
        count = torch.zeros(size[dim], dtype=src.dtype, device=src.device).scatter_add_(0, index, ones)
        count.clamp_(1)
        count = _broadcast(count, res, dim)
        if res.is_floating_point():
            res.true_divide_(count)
        else:
            res.div_(count, rounding_mode="floor")
        return res
    if reduce in ("min", "max"):
        red = "amin" if reduce == "min" else "amax"
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_reduce_(dim, bidx, src, red, include_self=False)
    raise ValueError(reduce)
