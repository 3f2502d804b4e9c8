"""Pins for the two third-party behaviours the oracle RESTATES: torch_scatter's "empty segment -> 0" for
min / max and DGL's zero rows for nodes that are never reduced.  Neither library is installed in the authoring container;
when a box has them (e.g. a reference install under baseline/_ref, or site-packages), these tests execute the real thing
against the restatement.  They skip -- visibly -- where the libraries are absent."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)


def _graph(n=60, e=300, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, int(n * 0.8), (e,), generator=g)       # the last 20 % of the nodes have no in-edge
    return src, dst, torch.randn(n, 7, generator=g)


def test_real_torch_scatter_matches_the_restated_scatter():
    torch_scatter = pytest.importorskip("torch_scatter", reason="torch_scatter is not installed on this machine")
    from oracle import pna_oracle as O
    src, dst, x = _graph()
    n = x.size(0)
    msg = x[src]
    for red in ("sum", "mean", "min", "max"):
        real = torch_scatter.scatter(msg, dst, 0, None, n, reduce=red)
        assert torch.equal(real, O.scatter(msg, dst, n, red)), red
    iso = torch.bincount(dst, minlength=n) == 0
    assert iso.any() and torch_scatter.scatter(msg, dst, 0, None, n, reduce="min")[iso].abs().max() == 0


def test_real_pyg_degree_and_propagate_match_the_restatement():
    pyg = pytest.importorskip("torch_geometric", reason="torch_geometric is not installed on this machine")
    pytest.importorskip("torch_scatter", reason="torch_scatter is not installed on this machine")
    from torch_geometric.utils import degree
    from oracle import pna_oracle as O
    src, dst, x = _graph(seed=1)
    assert torch.equal(degree(dst, x.size(0), dtype=x.dtype), O.degree(dst, x.size(0), dtype=x.dtype))


def test_real_dgl_leaves_unreduced_nodes_zero():
    dgl = pytest.importorskip("dgl", reason="dgl is not installed on this machine")
    from oracle import pna_oracle as O
    src, dst, x = _graph(seed=2)
    n = x.size(0)
    g = dgl.graph((src, dst), num_nodes=n)
    g.ndata["h"] = x
    A, S = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
    deg = torch.bincount(dst, minlength=n).float()
    avg = {"log": torch.log(deg + 1).mean().item()}

    def reduce_func(nodes):          # models/dgl/pna_layer.py:189-194 with the reference's aggregators / scalers
        h = nodes.mailbox["m"]
        D = h.shape[-2]
        mean, mx, mn = h.mean(-2), h.max(-2)[0], h.min(-2)[0]
        std = torch.sqrt(torch.relu((h * h).mean(-2) - mean * mean) + 1e-5)
        hh = torch.cat([mean, mx, mn, std], 1)
        import numpy as np
        return {"h": torch.cat([hh, hh * (np.log(D + 1) / avg["log"]), hh * (avg["log"] / np.log(D + 1))], 1)}
    g.update_all(dgl.function.copy_u("h", "m"), reduce_func)
    real = g.ndata["h"]
    want = O.dgl_reduce(x[src], None, dst, n, A, S, avg)
    iso = deg == 0
    assert iso.any() and real[iso].abs().max() == 0                 # the zero fill the oracle restates
    torch.testing.assert_close(real, want, rtol=1e-6, atol=1e-6)
