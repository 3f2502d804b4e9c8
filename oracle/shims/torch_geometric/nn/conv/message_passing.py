"""Shim of torch_geometric.nn.conv.MessagePassing for flow='source_to_target', node_dim=0, aggr=None.

Upstream behaviour restated (torch_geometric/nn/conv/message_passing.py, 1.6):
  propagate(edge_index, size=None, **kwargs):
    __collect__: for every argument of `message` named `<name>_j` / `<name>_i` gather kwargs[<name>] along node_dim
                 with edge_index[0] (source j) / edge_index[1] (target i); other arguments are passed through;
                 index = edge_index[1], dim_size = number of target nodes.
    out = self.aggregate(self.message(...), index=index, dim_size=dim_size); out = self.update(out)  (identity)
"""
import inspect

import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        assert size is None
        params = list(inspect.signature(self.message).parameters)
        n = None
        args = {}
        for name in params:
            if name.endswith("_i") or name.endswith("_j"):
                data = kwargs[name[:-2]]
                n = data.size(self.node_dim)
                sel = edge_index[1] if name.endswith("_i") else edge_index[0]
                args[name] = data.index_select(self.node_dim, sel)
            else:
                args[name] = kwargs.get(name)
        out = self.message(**args)
        out = self.aggregate(out, index=edge_index[1], dim_size=n)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
