"""DGL-signature PNA layers (reference ``models/dgl/pna_layer.py``) on the sm_100a aggregation kernel.

Same constructors, same ``forward(g, h, e, snorm_n)`` / ``forward(g, h)``, same parameter names
(``towers.{t}.pretrans.fully_connected.{k}.linear``, ``...posttrans...``, ``towers.{t}.batchnorm_h``,
``mixing_network.linear``; ``posttrans`` / ``batchnorm_h`` for the simple layer).  ``g`` is duck-typed (graph.py).
DGL's ``apply_edges`` + ``update_all`` with a Python reduce UDF per in-degree bucket (pna_layer.py:61-64,202) is
replaced by ONE kernel call over all towers; in-degree-0 nodes keep DGL's zero rows (PNA_FLAG_ZERO_ISOLATED).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import padding as pad
from .aggregate import pna_aggregate, row_scales
from .linear import compact_path_ok
from .csr import tensor_version
from .graph import graph_csr
from .nn_blocks import FCLayer, MLP

_AGGRS = ("mean", "sum", "max", "min", "std", "var")       # models/dgl/aggregators.py:50-52 minus moment3/4/5
_SCALERS = ("identity", "amplification", "attenuation")     # models/dgl/scalers.py:22


def _split(names, allowed, what):
    names = names.split() if isinstance(names, str) else list(names)
    for n in names:
        if n not in allowed:
            raise KeyError(f"{what} {n!r} is not available on the CUDA path (supported: {allowed})")
    return names


def _avg(avg_d) -> dict:
    return {k: float(v) for k, v in avg_d.items()}


class PNATower(nn.Module):
    """Parameters of one tower (pna_layer.py:17-33).  The aggregation itself runs once for all towers in PNALayer."""

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, pretrans_layers,
                 posttrans_layers, edge_features, edge_dim):
        super().__init__()
        self.dropout, self.graph_norm, self.batch_norm, self.edge_features = dropout, graph_norm, batch_norm, edge_features
        self.in_dim = in_dim
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.pretrans = MLP(in_size=2 * in_dim + (edge_dim if edge_features else 0), hidden_size=in_dim, out_size=in_dim,
                            layers=pretrans_layers, mid_activation="relu", last_activation="none")
        self.posttrans = MLP(in_size=(len(aggregators) * len(scalers) + 1) * in_dim, hidden_size=out_dim, out_size=out_dim,
                             layers=posttrans_layers, mid_activation="relu", last_activation="none")

    def finish(self, h_cat, snorm_n, blocks=None, fp=None):
        """posttrans -> graph norm -> batch norm -> dropout (pna_layer.py:67-76).  h_cat may carry padded column blocks
        (width fp instead of in_dim): the first posttrans Linear then gets zero weight columns at the pad positions."""
        if fp is not None and fp != self.in_dim:
            w0 = pad.expand_weight_cols(self.posttrans.fully_connected[0].linear.weight, blocks, self.in_dim, fp)
            h = self.posttrans(h_cat, first_weight=w0)
        else:
            h = self.posttrans(h_cat)
        if self.graph_norm:
            h = h * snorm_n
        if self.batch_norm:
            h = self.batchnorm_h(h)
        return F.dropout(h, self.dropout, training=self.training)


class PNALayer(nn.Module):
    """reference pna_layer.py:79-148."""

    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, graph_norm, batch_norm, towers=1,
                 pretrans_layers=1, posttrans_layers=1, divide_input=True, residual=False, edge_features=False, edge_dim=0):
        super().__init__()
        assert (not divide_input) or in_dim % towers == 0, "if divide_input is set the number of towers has to divide in_dim"
        assert out_dim % towers == 0, "the number of towers has to divide the out_dim"
        assert avg_d is not None
        self.aggregators = _split(aggregators, _AGGRS, "aggregator")
        self.scalers = _split(scalers, _SCALERS, "scaler")
        self.avg_d = _avg(avg_d)
        self.divide_input = divide_input
        self.input_tower = in_dim // towers if divide_input else in_dim
        self.output_tower = out_dim // towers
        self.in_dim, self.out_dim = in_dim, out_dim
        self.edge_features = edge_features
        self.residual = residual and in_dim == out_dim
        self.towers = nn.ModuleList([
            PNATower(in_dim=self.input_tower, out_dim=self.output_tower, aggregators=self.aggregators, scalers=self.scalers,
                     avg_d=avg_d, pretrans_layers=pretrans_layers, posttrans_layers=posttrans_layers, batch_norm=batch_norm,
                     dropout=dropout, graph_norm=graph_norm, edge_features=edge_features, edge_dim=edge_dim)
            for _ in range(towers)])
        self.mixing_network = FCLayer(out_dim, out_dim, activation="LeakyReLU")

    def _tower_input(self, h, t):
        it = self.input_tower
        return h[:, t * it:(t + 1) * it] if self.divide_input else h

    def _affine_terms(self, h, fp):
        """pretrans(cat[src h, dst h]) = W_s h_src + W_d h_dst + b (pna_layer.py:35-40): V = h W_s^T + b, U = h W_d^T,
        each tower block padded to fp columns with zero weight rows.  Both come out of ONE GEMM against the packed weight
        [W_d ; W_s] (block-diagonal per tower with divide_input), which is rebuilt only when a parameter changed (without
        autograd; with autograd it is part of the graph and rebuilt every call)."""
        it = self.input_tower
        lins = [tw.pretrans.fully_connected[0].linear for tw in self.towers]
        params = [p_ for l in lins for p_ in (l.weight, l.bias)]
        key = (fp, tuple(tensor_version(p_) for p_ in params), tuple(p_.data_ptr() for p_ in params))
        cache = not (torch.is_grad_enabled() and any(p_.requires_grad for p_ in params))
        hit = getattr(self, "_uv_pack", None)
        if cache and hit is not None and hit[0] == key:
            w_uv, b_uv = hit[1]
        else:
            Ws = [pad.expand_weight_rows(l.weight[:, :it], it, fp) for l in lins]
            Wd = [pad.expand_weight_rows(l.weight[:, it:2 * it], it, fp) for l in lins]
            b = torch.cat([F.pad(l.bias, (0, fp - it)) for l in lins])
            if self.divide_input and len(lins) > 1:
                w_uv = torch.cat([torch.block_diag(*Wd), torch.block_diag(*Ws)], 0)
            else:
                w_uv = torch.cat(Wd + Ws, 0)
            b_uv = torch.cat([torch.zeros_like(b), b])
            if cache:
                self._uv_pack = (key, (w_uv, b_uv))
        uv = torch.addmm(b_uv, h, w_uv.t())
        half = uv.size(1) // 2
        return uv[:, :half], uv[:, half:]

    def _edge_messages(self, csr, h, e):
        src, dst = csr.col.long(), csr.dst_of_slot
        ef = e.index_select(0, csr.perm.long()) if self.edge_features else None
        msgs = []
        for t, tw in enumerate(self.towers):
            ht = self._tower_input(h, t)
            parts = [ht.index_select(0, src), ht.index_select(0, dst)] + ([ef] if ef is not None else [])
            msgs.append(tw.pretrans(torch.cat(parts, dim=1)))
        return torch.cat(msgs, dim=1)

    def forward(self, g, h, e, snorm_n):
        h_in = h
        csr = graph_csr(g, h.device)
        T, it = len(self.towers), self.input_tower
        fp = pad.padded_width(it, h.dtype)
        if fp == it:
            h_self = h
        elif self.divide_input:
            h_self = pad.pad_blocks(h, T, it, fp)
        else:
            h_self = pad.pad_cols(h, fp)
        common = dict(towers=T, self_feat=h_self, self_divided=self.divide_input, zero_isolated=True, relu_var=True)
        if not self.edge_features and self.towers[0].pretrans.is_single_affine():
            U, V = self._affine_terms(h, fp)
            agg = pna_aggregate(V, csr, self.aggregators, self.scalers, self.avg_d, row_bias=U, **common)
        else:
            msgs = pad.pad_blocks(self._edge_messages(csr, h, e), T, it, fp)
            agg = pna_aggregate(msgs, csr, self.aggregators, self.scalers, self.avg_d, messages_in_csr_order=True, **common)
        agg = agg.view(h.size(0), T, -1)                                  # [N, T, (1 + S*A) * fp] = cat([h_t, reduced])
        blocks = 1 + len(self.aggregators) * len(self.scalers)
        h_cat = torch.cat([tw.finish(agg[:, t], snorm_n, blocks, fp) for t, tw in enumerate(self.towers)], dim=1)
        h_out = self.mixing_network(h_cat)
        if self.residual:
            h_out = h_in + h_out
        return h_out

    def __repr__(self):
        return f"{self.__class__.__name__}(in_channels={self.in_dim}, out_channels={self.out_dim})"


class PNASimpleLayer(nn.Module):
    """reference pna_layer.py:151-219: aggregate the neighbours' h directly, posttrans, BN, ReLU, residual, dropout."""

    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, batch_norm, residual, posttrans_layers=1):
        super().__init__()
        self.aggregators = _split(aggregators, _AGGRS, "aggregator")
        self.scalers = _split(scalers, _SCALERS, "scaler")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.dropout, self.batch_norm, self.residual = dropout, batch_norm, residual
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.posttrans = MLP(in_size=(len(self.aggregators) * len(self.scalers)) * in_dim, hidden_size=out_dim, out_size=out_dim,
                             layers=posttrans_layers, mid_activation="relu", last_activation="none")
        self.avg_d = _avg(avg_d)

    def aggregate_only(self, g, h):
        return pna_aggregate(h, graph_csr(g, h.device), self.aggregators, self.scalers, self.avg_d, zero_isolated=True, relu_var=True)

    def forward(self, g, h):
        h_in = h
        fp = pad.padded_width(self.in_dim, h.dtype)
        csr = graph_csr(g, h.device)
        blocks = len(self.aggregators) * len(self.scalers)
        w0 = self.posttrans.fully_connected[0].linear.weight
        if fp != self.in_dim:   # odd width: 128-bit path on zero-padded rows, padding absorbed by the first posttrans Linear
            w0 = pad.expand_weight_cols(w0, blocks, self.in_dim, fp)
        hp = pad.pad_cols(h, fp)
        if w0.dtype == torch.float32 and compact_path_ok(h, len(self.aggregators) * fp, w0.size(0), len(self.scalers)):
            # compact post path: identity-scaled aggregate, the scaled copies are formed inside the tensor-core linear
            agg = pna_aggregate(hp, csr, self.aggregators, ["identity"], self.avg_d, zero_isolated=True, relu_var=True)
            h = self.posttrans(agg, first_weight=w0, first_row_scale=row_scales(csr, self.scalers, self.avg_d))
        else:
            agg = pna_aggregate(hp, csr, self.aggregators, self.scalers, self.avg_d, zero_isolated=True, relu_var=True)
            h = self.posttrans(agg, first_weight=w0)
        if self.batch_norm:
            h = self.batchnorm_h(h)
        h = F.relu(h)
        if self.residual:
            h = h_in + h
        return F.dropout(h, self.dropout, training=self.training)

    def __repr__(self):
        return f"{self.__class__.__name__}(in_channels={self.in_dim}, out_channels={self.out_dim})"
