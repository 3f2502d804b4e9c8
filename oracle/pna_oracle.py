"""CPU ORACLE for the PNA layer forward -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  Nothing under ``pna_b200/`` imports it; the product has no CPU path.

It restates, op for op with plain torch CPU ops, the reference's PyG path
(``models/pytorch_geometric/{pna,aggregators,scalers}.py``) and its DGL path (``models/dgl/*``), including the
behaviour of the third-party calls those files make, which are NOT part of /root/reference and are not pinned there
(no PyG pin anywhere; conda env pins dgl 0.4.2, realworld_benchmark/environment_gpu.yml:20):

  torch_scatter.scatter(src, index, 0, None, dim_size, reduce=...)     aggregators.py:10,14,18,22
      sum : zeros(dim_size).scatter_add_(0, index_broadcast, src)
      mean: sum / count.clamp_(1)   (count = scatter_add of ones; true division)
      min/max: reduce over the rows of each segment; segments with no row give 0 (the reducer's init value is
               masked to 0 when no `out` is passed)
  torch_geometric MessagePassing.propagate(edge_index, x=...)          pna.py:129,236
      x_j = x.index_select(0, edge_index[0]); x_i = x.index_select(0, edge_index[1]);
      aggregate(message(...), index=edge_index[1], dim_size=N); update = identity
  torch_geometric.utils.degree(index, N, dtype)                         pna.py:157,247
      zeros(N, dtype).scatter_add_(0, index, ones)
  dgl 0.4 update_all(message, reduce)                                   models/dgl/pna_layer.py:64,202
      nodes are bucketed by in-degree D>0; reduce sees a mailbox [n_D, D, F] in edge-id order;
      nodes with in-degree 0 are not reduced and get the zero initialiser.

PARITY PINNING.  The reference ships no tests and no golden vectors, and torch_geometric /
torch_scatter / dgl cannot be installed here.  The oracle is pinned by (see tests/test_oracle.py and
oracle/gen_golden.py):
  (1) the REAL reference files models/pytorch_geometric/pna.py, aggregators.py, scalers.py and models/dgl/*.py,
      executed in this container over minimal restatements of the missing third-party modules (oracle/shims/),
      whose outputs are committed as tests/golden/*.pt;
  (2) the REAL dense reference models/pytorch/pna/{aggregators,scalers}.py (imports here unmodified), known answers
      K1 of SURVEY.md section 8c;
  (3) the reference's numpy neighbourhood reducers multitask_benchmark/datasets_generation/graph_algorithms.py:61-114;
  (4) analytic cases (in-degree 0 and 1).
The two third-party behaviours marked "unverified" in SURVEY.md 8c (torch_scatter empty-segment -> 0; DGL zero fill
of isolated nodes) are restated from the upstream sources' documented behaviour and remain unexecuted here.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor
from torch.nn import Linear, Module, ModuleList, ReLU, Sequential

EPS = 1e-5


# ---- third-party behaviours restated ------------------------------------------------------------------------------
def _broadcast_index(index: Tensor, src: Tensor) -> Tensor:
    view = [index.numel()] + [1] * (src.dim() - 1)
    return index.view(view).expand_as(src)


def scatter(src: Tensor, index: Tensor, dim_size: int, reduce: str) -> Tensor:
    """torch_scatter.scatter(src, index, 0, None, dim_size, reduce) along dim 0."""
    size = [dim_size] + list(src.shape[1:])
    idx = _broadcast_index(index, src)
    if reduce in ("sum", "add"):
        return torch.zeros(size, dtype=src.dtype).scatter_add_(0, idx, src)
    if reduce == "mean":
        out = torch.zeros(size, dtype=src.dtype).scatter_add_(0, idx, src)
        count = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
        count.clamp_(1)
        return out.true_divide_(count.view([dim_size] + [1] * (src.dim() - 1)))
    if reduce in ("min", "max"):
        red = "amin" if reduce == "min" else "amax"
        return torch.zeros(size, dtype=src.dtype).scatter_reduce_(0, idx, src, red, include_self=False)
    raise KeyError(reduce)


def degree(index: Tensor, num_nodes: int, dtype=torch.float32) -> Tensor:
    return torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, index, torch.ones(index.numel(), dtype=dtype))


# ---- reference aggregators / scalers (PyG flavour) ------------------------------------------------------------------
def aggregate_sum(src, index, dim_size):      # aggregators.py:9-10
    return scatter(src, index, dim_size, "sum")


def aggregate_mean(src, index, dim_size):     # aggregators.py:13-14
    return scatter(src, index, dim_size, "mean")


def aggregate_min(src, index, dim_size):      # aggregators.py:17-18
    return scatter(src, index, dim_size, "min")


def aggregate_max(src, index, dim_size):      # aggregators.py:21-22
    return scatter(src, index, dim_size, "max")


def aggregate_var(src, index, dim_size):      # aggregators.py:25-28
    mean = aggregate_mean(src, index, dim_size)
    mean_squares = aggregate_mean(src * src, index, dim_size)
    return mean_squares - mean * mean


def aggregate_std(src, index, dim_size):      # aggregators.py:31-32
    return torch.sqrt(torch.relu(aggregate_var(src, index, dim_size)) + EPS)


AGGREGATORS = {"sum": aggregate_sum, "mean": aggregate_mean, "min": aggregate_min, "max": aggregate_max,
               "var": aggregate_var, "std": aggregate_std}


def scale_identity(src, deg, avg_deg):         # scalers.py:8-9
    return src


def scale_amplification(src, deg, avg_deg):    # scalers.py:12-13
    return src * (torch.log(deg + 1) / avg_deg["log"])


def scale_attenuation(src, deg, avg_deg):      # scalers.py:16-19
    scale = avg_deg["log"] / torch.log(deg + 1)
    scale[deg == 0] = 1
    return src * scale


def scale_linear(src, deg, avg_deg):           # scalers.py:22-23
    return src * (deg / avg_deg["lin"])


def scale_inverse_linear(src, deg, avg_deg):   # scalers.py:26-29
    scale = avg_deg["lin"] / deg
    scale[deg == 0] = 1
    return src * scale


SCALERS = {"identity": scale_identity, "amplification": scale_amplification, "attenuation": scale_attenuation,
           "linear": scale_linear, "inverse_linear": scale_inverse_linear}


def avg_deg_from_histogram(deg: Tensor) -> Dict[str, float]:
    """pna.py:79-86."""
    deg = deg.to(torch.float)
    total_no_vertices = deg.sum()
    bin_degrees = torch.arange(len(deg))
    return {
        "lin": ((bin_degrees * deg).sum() / total_no_vertices).item(),
        "log": (((bin_degrees + 1).log() * deg).sum() / total_no_vertices).item(),
        "exp": ((bin_degrees.exp() * deg).sum() / total_no_vertices).item(),
    }


def pyg_aggregate(inputs: Tensor, index: Tensor, dim_size: int, aggregators: Sequence[str], scalers: Sequence[str],
                  avg_deg: Dict[str, float]) -> Tensor:
    """PNAConv.aggregate / PNAConvSimple.aggregate (pna.py:152-159, :242-249).  inputs: [E, F] or [E, T, F]."""
    outs = [AGGREGATORS[a](inputs, index, dim_size) for a in aggregators]
    out = torch.cat(outs, dim=-1)
    deg = degree(index, dim_size, dtype=inputs.dtype).view([-1] + [1] * (inputs.dim() - 1))
    outs = [SCALERS[s](out, deg, avg_deg) for s in scalers]
    return torch.cat(outs, dim=-1)


def simple_propagate(x: Tensor, edge_index: Tensor, aggregators, scalers, avg_deg) -> Tensor:
    """PNAConvSimple.propagate: message = x_j (pna.py:236-240) followed by aggregate."""
    x_j = x.index_select(0, edge_index[0])
    return pyg_aggregate(x_j, edge_index[1], x.size(0), aggregators, scalers, avg_deg)


def _reset(nn: Module) -> None:
    for m in nn.modules():
        if m is not nn and hasattr(m, "reset_parameters"):
            m.reset_parameters()


class PNAConvSimpleOracle(Module):
    """pna.py:167-254 on CPU tensors."""

    def __init__(self, in_channels, out_channels, aggregators: List[str], scalers: List[str], deg: Tensor,
                 post_layers: int = 1):
        super().__init__()
        self.aggregators, self.scalers = list(aggregators), list(scalers)
        self.F_in, self.F_out = in_channels, out_channels
        self.avg_deg = avg_deg_from_histogram(deg)
        modules = [Linear(len(aggregators) * len(scalers) * self.F_in, self.F_out)]
        for _ in range(post_layers - 1):
            modules += [ReLU(), Linear(self.F_out, self.F_out)]
        self.post_nn = Sequential(*modules)

    def propagate(self, x, edge_index):
        return simple_propagate(x, edge_index, self.aggregators, self.scalers, self.avg_deg)

    def forward(self, x, edge_index, edge_attr=None):
        return self.post_nn(self.propagate(x, edge_index))


class PNAConvOracle(Module):
    """pna.py:17-164 on CPU tensors."""

    def __init__(self, in_channels, out_channels, aggregators, scalers, deg, edge_dim=None, towers=1, pre_layers=1,
                 post_layers=1, divide_input=False):
        super().__init__()
        self.aggregators, self.scalers = list(aggregators), list(scalers)
        self.edge_dim, self.towers, self.divide_input = edge_dim, towers, divide_input
        self.F_in = in_channels // towers if divide_input else in_channels
        self.F_out = out_channels // towers
        self.avg_deg = avg_deg_from_histogram(deg)
        if edge_dim is not None:
            self.edge_encoder = Linear(edge_dim, self.F_in)
        self.pre_nns, self.post_nns = ModuleList(), ModuleList()
        for _ in range(towers):
            modules = [Linear((3 if edge_dim else 2) * self.F_in, self.F_in)]
            for _ in range(pre_layers - 1):
                modules += [ReLU(), Linear(self.F_in, self.F_in)]
            self.pre_nns.append(Sequential(*modules))
            modules = [Linear((len(aggregators) * len(scalers) + 1) * self.F_in, self.F_out)]
            for _ in range(post_layers - 1):
                modules += [ReLU(), Linear(self.F_out, self.F_out)]
            self.post_nns.append(Sequential(*modules))
        self.lin = Linear(out_channels, out_channels)

    def message(self, x_i, x_j, edge_attr):                     # pna.py:137-150
        if edge_attr is not None:
            edge_attr = self.edge_encoder(edge_attr)
            edge_attr = edge_attr.view(-1, 1, self.F_in).repeat(1, self.towers, 1)
            h = torch.cat([x_i, x_j, edge_attr], dim=-1)
        else:
            h = torch.cat([x_i, x_j], dim=-1)
        hs = [nn(h[:, i]) for i, nn in enumerate(self.pre_nns)]
        return torch.stack(hs, dim=1)

    def propagate(self, x, edge_index, edge_attr=None):
        x_j = x.index_select(0, edge_index[0])
        x_i = x.index_select(0, edge_index[1])
        msg = self.message(x_i, x_j, edge_attr)
        return pyg_aggregate(msg, edge_index[1], x.size(0), self.aggregators, self.scalers, self.avg_deg)

    def forward(self, x, edge_index, edge_attr=None):          # pna.py:120-135
        if self.divide_input:
            x = x.view(-1, self.towers, self.F_in)
        else:
            x = x.view(-1, 1, self.F_in).repeat(1, self.towers, 1)
        out = self.propagate(x, edge_index, edge_attr)
        out = torch.cat([x, out], dim=-1)
        outs = [nn(out[:, i]) for i, nn in enumerate(self.post_nns)]
        out = torch.cat(outs, dim=1)
        return self.lin(out)


# ---- DGL flavour ----------------------------------------------------------------------------------------------------
def dgl_reduce(messages: Tensor, src_unused: Optional[Tensor], dst: Tensor, num_nodes: int, aggregators: Sequence[str],
               scalers: Sequence[str], avg_d: Dict[str, float]) -> Tensor:
    """update_all(message, reduce_func) of models/dgl/pna_layer.py:45-50,189-194 with DGL 0.4 degree bucketing.

    messages: [E, F] per-edge messages in edge-id order.  Returns [N, S*A*F]; in-degree-0 rows are 0.
    """
    import numpy as np
    F = messages.size(1)
    A, S = len(aggregators), len(scalers)
    out = torch.zeros((num_nodes, S * A * F), dtype=messages.dtype)
    deg = torch.bincount(dst, minlength=num_nodes)
    order = torch.sort(dst, stable=True).indices          # mailbox rows in edge-id order per destination
    start = torch.cumsum(deg, 0) - deg
    for D in torch.unique(deg).tolist():
        if D == 0:
            continue
        nodes = (deg == D).nonzero().flatten()
        slots = (start[nodes].unsqueeze(1) + torch.arange(D).unsqueeze(0))          # [n, D]
        h = messages[order[slots]]                                                    # mailbox [n, D, F]
        cols = []
        for a in aggregators:                                                         # models/dgl/aggregators.py:6-26
            if a == "mean":
                cols.append(torch.mean(h, dim=1))
            elif a == "max":
                cols.append(torch.max(h, dim=1)[0])
            elif a == "min":
                cols.append(torch.min(h, dim=1)[0])
            elif a == "sum":
                cols.append(torch.sum(h, dim=1))
            elif a in ("var", "std"):
                var = torch.relu(torch.mean(h * h, dim=-2) - torch.mean(h, dim=-2) * torch.mean(h, dim=-2))
                cols.append(var if a == "var" else torch.sqrt(var + EPS))
            else:
                raise KeyError(a)
        hh = torch.cat(cols, dim=1)
        sc = []
        for s in scalers:                                                             # models/dgl/scalers.py:8-19
            if s == "identity":
                sc.append(hh)
            elif s == "amplification":
                sc.append(hh * (np.log(D + 1) / avg_d["log"]))
            elif s == "attenuation":
                sc.append(hh * (avg_d["log"] / np.log(D + 1)))
            else:
                raise KeyError(s)
        out[nodes] = torch.cat(sc, dim=1)
    return out
