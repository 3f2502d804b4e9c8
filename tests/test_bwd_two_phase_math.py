"""The algebra behind the two-phase backward (pna_aggregate_bwd_coef + a 'sum' aggregation over the transposed graph +
pna_aggregate_bwd_combine), restated in torch float64 and checked against the reference's autograd (CPU oracle).

For a destination row i with messages m_s = gathered[col[s]] + row_bias[i] the gradient of one message is
    grad_m(s) = c0_i + c1_i * m_s + [s == argmin_i] gmin_i + [s == argmax_i] gmax_i            (csrc/pna_aggregate_bwd.cu)
so that, per SOURCE row j,
    grad_gathered[j] = sum_{i <- j} (c0_i + c1_i * row_bias[i])  +  gathered[j] * sum_{i <- j} c1_i  +  routed min / max terms
and  grad_row_bias[i] = deg_i * c0_i + c1_i * sum_s m_s + gmin_i + gmax_i.
The two sums over the out-edges of j are a plain 'sum' aggregation of the rows [c0' | c1] over the transposed graph: no atomics
except the one add per (row, feature) that routes min and max.
"""
import math

import pytest
import torch

from oracle import pna_oracle as O

AGGRS = ["sum", "mean", "min", "max", "var", "std"]
SCALERS = ["identity", "amplification", "attenuation", "linear", "inverse_linear"]


def _scales(deg: int, avg):
    lg = math.log(deg + 1.0)
    return {"identity": 1.0, "amplification": lg / avg["log"], "attenuation": avg["log"] / lg if deg else 1.0,
            "linear": deg / avg["lin"], "inverse_linear": avg["lin"] / deg if deg else 1.0}


def two_phase_backward(x, bias, src, dst, n, grad_out, aggrs, scalers, avg):
    """Phase 1 per destination row, phase 2 per source row; everything in the dtype of x."""
    f = x.size(1)
    order = torch.sort(dst, stable=True).indices
    col, dst_s = src[order], dst[order]
    rowptr = torch.zeros(n + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    coef = torch.zeros(n, 2 * f, dtype=x.dtype)            # rows without in-edges are never read in phase 2
    gg = torch.zeros(x.size(0), f, dtype=x.dtype)
    gb = torch.zeros(n, f, dtype=x.dtype)
    A = len(aggrs)
    for i in range(n):
        beg, end = int(rowptr[i]), int(rowptr[i + 1])
        deg = end - beg
        if deg == 0:
            continue
        b = bias[i] if bias is not None else torch.zeros(f, dtype=x.dtype)
        m = x[col[beg:end]] + b
        s, sq = m.sum(0), (m * m).sum(0)
        mn, amn = m.min(0)
        mx, amx = m.max(0)
        # first slot attaining the extremum (torch.min returns an unspecified one among ties)
        amn = torch.stack([(m[:, j] == mn[j]).nonzero()[0, 0] for j in range(f)])
        amx = torch.stack([(m[:, j] == mx[j]).nonzero()[0, 0] for j in range(f)])
        mean = s / deg
        var = sq / deg - mean * mean
        sd = torch.sqrt(torch.relu(var) + 1e-5)
        sc = _scales(deg, avg)
        c0 = torch.zeros(f, dtype=x.dtype); c1 = torch.zeros(f, dtype=x.dtype)
        gmin = torch.zeros(f, dtype=x.dtype); gmax = torch.zeros(f, dtype=x.dtype)
        for a, name in enumerate(aggrs):
            g = sum(sc[sn] * grad_out[i, (k * A + a) * f:(k * A + a + 1) * f] for k, sn in enumerate(scalers))
            if name == "sum":
                c0 += g
            elif name == "mean":
                c0 += g / deg
            elif name == "min":
                gmin += g
            elif name == "max":
                gmax += g
            elif name == "var":
                t = 2.0 * g / deg
                c1 += t; c0 -= t * mean
            else:
                t = torch.where(var > 0, g / (deg * sd), torch.zeros_like(g))
                c1 += t; c0 -= t * mean
        coef[i, :f] = c0 + c1 * b
        coef[i, f:] = c1
        cols = torch.arange(f)
        gg.index_put_((col[beg + amn], cols), gmin, accumulate=True)
        gg.index_put_((col[beg + amx], cols), gmax, accumulate=True)
        gb[i] = deg * c0 + c1 * s + gmin + gmax
    sums = torch.zeros(x.size(0), 2 * f, dtype=x.dtype).index_add_(0, col, coef[dst_s])     # 'sum' over the transposed graph
    gg += sums[:, :f] + x * sums[:, f:]
    return gg, gb


@pytest.mark.parametrize("with_bias", [False, True])
def test_two_phase_backward_equals_reference_autograd(with_bias):
    g = torch.Generator().manual_seed(5)
    n, e, f = 40, 260, 5
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - 6, (e,), generator=g)              # the last rows stay isolated
    dst[:90] = 3                                                  # one large row
    x = torch.randn(n, f, generator=g, dtype=torch.float64)
    bias = torch.randn(n, f, generator=g, dtype=torch.float64) if with_bias else None
    deg_hist = torch.bincount(torch.bincount(dst, minlength=n))
    avg = O.avg_deg_from_histogram(deg_hist)
    w = torch.randn(n, len(AGGRS) * len(SCALERS) * f, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if with_bias else None
    msg = xr[src] + (br[dst] if with_bias else 0.0)
    out = O.pyg_aggregate(msg, dst, n, AGGRS, SCALERS, avg)
    (out * w).sum().backward()
    gg, gb = two_phase_backward(x, bias, src, dst, n, w, AGGRS, SCALERS, avg)
    torch.testing.assert_close(gg, xr.grad, rtol=1e-9, atol=1e-9)
    if with_bias:
        torch.testing.assert_close(gb, br.grad, rtol=1e-9, atol=1e-9)
