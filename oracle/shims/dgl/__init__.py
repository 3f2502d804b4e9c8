"""Shim of dgl 0.4.x -- only what reference models/dgl/pna_layer.py uses: g.ndata / g.edata dictionaries,
g.apply_edges(func), g.update_all(message_func, reduce_func) and dgl.function.copy_u.

Upstream behaviour restated (dgl 0.4 python/dgl/graph.py + runtime/scheduler.py, degree-bucketing executor):
  apply_edges(func): func(EdgeBatch) with .src / .dst / .data dicts gathered per edge (edge-id order); the returned
      dict is written to edata.
  update_all(message_func, reduce_func): messages are computed for all edges; destination nodes are bucketed by
      in-degree D; reduce_func(NodeBatch) sees mailbox[field] of shape [n_D, D, ...] (messages of a node in edge-id
      order) and returns per-node features; nodes with in-degree 0 are skipped and their rows of a NEWLY created
      output column come from the frame initialiser (zeros).
"""
import torch

from . import function  # noqa: F401


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _NodeBatch:
    def __init__(self, data, mailbox):
        self.data, self.mailbox = data, mailbox


class DGLGraph:
    def __init__(self, src, dst, num_nodes):
        self._src, self._dst, self._n = src.long(), dst.long(), int(num_nodes)
        self.ndata, self.edata = {}, {}

    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def _edge_batch(self):
        return _EdgeBatch({k: v[self._src] for k, v in self.ndata.items()},
                          {k: v[self._dst] for k, v in self.ndata.items()}, dict(self.edata))

    def apply_edges(self, func):
        self.edata.update(func(self._edge_batch()))

    def update_all(self, message_func, reduce_func):
        msgs = message_func(self._edge_batch())
        deg = self.in_degrees()
        order = torch.sort(self._dst, stable=True).indices
        start = torch.cumsum(deg, 0) - deg
        results = {}
        for D in torch.unique(deg).tolist():
            if D == 0:
                continue
            nodes = (deg == D).nonzero().flatten()
            slots = start[nodes].unsqueeze(1) + torch.arange(D).unsqueeze(0)
            mailbox = {k: v[order[slots]] for k, v in msgs.items()}
            red = reduce_func(_NodeBatch({k: v[nodes] for k, v in self.ndata.items()}, mailbox))
            for k, v in red.items():
                if k not in results:
                    results[k] = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype)
                results[k][nodes] = v
        self.ndata.update(results)
