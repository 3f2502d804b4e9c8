"""The backward kernels (csrc/pna_aggregate_bwd.cu) executed on the HOST, thread by thread (tests/emu): their index arithmetic,
split-row path and the coefficient mode (pna_aggregate_bwd_coef / pna_aggregate_bwd_combine) against the reference's autograd
(CPU oracle) and against each other.  The kernels have no intra-block communication, so sequential execution is faithful;
what this cannot see (memory ordering between concurrent warps, the forward kernels that sum the coefficient rows on the
GPU) is covered by the -m gpu tests."""
import ctypes as C
import shutil

import pytest
import torch

from oracle import pna_oracle as O
from pna_b200 import _lib

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

AGGRS = ["sum", "mean", "min", "max", "var", "std"]
SCALERS = ["identity", "amplification", "attenuation", "linear", "inverse_linear"]
SPLIT, CHUNK = 16, 8


@pytest.fixture(scope="module")
def emu():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "build_emu.py"))
    build_emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_emu)
    try:
        L = C.CDLL(build_emu.build("pna_aggregate_bwd.cu"))
    except Exception as exc:            # no CUDA headers on this machine
        pytest.skip(f"emulation library did not build: {exc}")
    L.emu_last_error.restype = C.c_char_p
    L.pna_aggregate_bwd.argtypes = [C.POINTER(_lib.AggStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    L.pna_aggregate_bwd_coef.argtypes = [C.POINTER(_lib.AggStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    L.pna_aggregate_bwd_combine.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                            C.c_int64, C.c_int32, C.c_void_p]
    return L


def host_csr(src, dst, n):
    """Destination-sorted CSR + the split-row tables, as pna_csr_build lays them out (include/pna_b200.h)."""
    order = torch.sort(dst, stable=True).indices
    col = src[order].to(torch.int32).contiguous()
    deg = torch.bincount(dst, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    hubs, chunks = [], []
    for r in (deg >= SPLIT).nonzero().flatten().tolist():
        nch = (int(deg[r]) + CHUNK - 1) // CHUNK
        hubs.append([r, len(chunks), nch, int(deg[r])])
        chunks += [[len(hubs) - 1, j] for j in range(nch)]
    hub_info = torch.tensor(hubs, dtype=torch.int32).reshape(-1, 4).contiguous()
    chunk_items = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).contiguous()
    return order, rowptr, col, hub_info, chunk_items


def descriptor(x, bias, rowptr, col, hub_info, chunk_items, n, aggrs, scalers, avg, towers, scratch):
    na, ac = _lib.pack_codes(aggrs, _lib.AGGR_CODES, "aggregator")
    ns, sc = _lib.pack_codes(scalers, _lib.SCALER_CODES, "scaler")
    f = x.size(1)
    return _lib.AggStruct(
        gathered=x.data_ptr(), ld_gathered=x.stride(0), rowptr=rowptr.data_ptr(), col=col.data_ptr(),
        row_bias=None if bias is None else bias.data_ptr(), ld_row_bias=0 if bias is None else bias.stride(0),
        n_rows=n, n_feat=f, n_towers=towers, dtype=_lib.PNA_F32 if x.dtype == torch.float32 else _lib.PNA_BF16,
        n_aggr=na, aggr_codes=ac, n_scalers=ns, scaler_codes=sc, avg_log=float(avg["log"]), avg_lin=float(avg["lin"]),
        split_threshold=SPLIT, chunk_edges=CHUNK, hub_info=hub_info.data_ptr() if hub_info.numel() else None,
        chunk_items=chunk_items.data_ptr() if chunk_items.numel() else None, n_hubs=hub_info.size(0), n_chunks=chunk_items.size(0),
        hub_partials=scratch.data_ptr())


def run_both(emu, x, bias, src, dst, n, w, aggrs, scalers, avg, towers=1):
    """(grad_gathered, grad_row_bias) of the one-call path and of the coefficient path, both on the host."""
    f, n_src = x.size(1), x.size(0)
    order, rowptr, col, hub_info, chunk_items = host_csr(src, dst, n)
    scratch = torch.zeros(((chunk_items.size(0) + hub_info.size(0)) * 6 + 1, f))
    d = descriptor(x, bias, rowptr, col, hub_info, chunk_items, n, aggrs, scalers, avg, towers, scratch)
    w = w.to(x.dtype).contiguous()
    gg1, gb1 = torch.zeros(n_src, f), torch.full((n, f), float("nan"))
    rc = emu.pna_aggregate_bwd(C.byref(d), w.data_ptr(), w.stride(0), gg1.data_ptr(), f, gb1.data_ptr(), f, None)
    assert rc == 0, emu.emu_last_error()
    fp = (f + 3) // 4 * 4
    coef = torch.full((n, 2 * fp), float("nan"))                  # rows without in-edges and pad columns stay NaN: never read
    gg2, gb2 = torch.zeros(n_src, f), torch.full((n, f), float("nan"))
    rc = emu.pna_aggregate_bwd_coef(C.byref(d), w.data_ptr(), w.stride(0), coef.data_ptr(), 2 * fp, fp, gg2.data_ptr(), f,
                                    gb2.data_ptr(), f, None)
    assert rc == 0, emu.emu_last_error()
    has_in = torch.bincount(dst, minlength=n) > 0
    assert torch.isfinite(coef[has_in][:, :f]).all() and torch.isfinite(coef[has_in][:, fp:fp + f]).all()
    assert torch.isnan(coef[~has_in]).all()
    # step 2 (on the GPU: pna_aggregate_fwd 'sum' over the transposed graph): the coefficient rows summed per source row
    dst_s = dst[order]
    sums = torch.zeros(n_src, 2 * fp)
    cz = torch.nan_to_num(coef, nan=0.0)
    sums.index_add_(0, col.long(), cz[dst_s])
    rc = emu.pna_aggregate_bwd_combine(sums.data_ptr(), 2 * fp, fp, x.data_ptr(), x.stride(0), d.dtype, gg2.data_ptr(), f, n_src, f, None)
    assert rc == 0, emu.emu_last_error()
    return (gg1, gb1), (gg2, gb2)


def reference_grads(x, bias, src, dst, n, w, aggrs, scalers, avg, towers=1):
    # fp32 like the reference runs it: where var ~ 0 the slope of sqrt(var + 1e-5) makes fp32 and fp64 gradients differ by 3e-3
    xr = x.float().clone().requires_grad_(True)
    br = bias.float().clone().requires_grad_(True) if bias is not None else None
    msg = xr[src] + (br[dst] if br is not None else 0.0)
    f = x.size(1)
    ft = f // towers
    outs = [O.pyg_aggregate(msg[:, t * ft:(t + 1) * ft], dst, n, aggrs, scalers, avg) for t in range(towers)]
    (torch.cat(outs, 1) * w.float()).sum().backward()
    return xr.grad.float(), None if br is None else br.grad.float()


def graph(n, e, big, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - 5, (e,), generator=g)          # the last rows are isolated
    dst[:big] = 2                                             # one row far above the split threshold (several chunks)
    dst[big:big + SPLIT] = 7                                  # one row exactly at it
    return src, dst, g


@pytest.mark.parametrize("f,towers,with_bias", [(12, 1, False), (12, 1, True), (10, 1, True), (16, 2, True), (40, 1, False), (3, 1, False),
                                                  (128, 1, False), (160, 1, True), (192, 4, True)])
def test_emulated_backward_matches_reference_autograd(emu, f, towers, with_bias):
    n, e = 60, 420
    src, dst, g = graph(n, e, 70, seed=f)
    x = torch.randn(n, f, generator=g)
    bias = torch.randn(n, f, generator=g) if with_bias else None
    avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(dst, minlength=n)))
    w = torch.randn(n, len(AGGRS) * len(SCALERS) * f, generator=g)
    (gg1, gb1), (gg2, gb2) = run_both(emu, x, bias, src, dst, n, w, AGGRS, SCALERS, avg, towers)
    want_x, want_b = reference_grads(x, bias, src, dst, n, w, AGGRS, SCALERS, avg, towers)
    # std has a slope of up to 158 at var ~ 0 and the sums run in fp32: the same bar as the GPU tests
    for got in (gg1, gg2):
        torch.testing.assert_close(got, want_x, rtol=1e-3, atol=5e-4)
    if with_bias:
        for got in (gb1, gb2):
            torch.testing.assert_close(got, want_b, rtol=1e-3, atol=2e-3)
    torch.testing.assert_close(gg2, gg1, rtol=1e-4, atol=2e-4)


def test_emulated_backward_routes_ties_to_the_first_slot_in_both_paths(emu):
    n, e, f = 50, 400, 8
    src, dst, g = graph(n, e, 90, seed=1)
    x = torch.randint(-2, 3, (n, f), generator=g).float()             # many ties
    avg = {"log": 1.3, "lin": 4.0}
    w = torch.ones(n, 2 * f)
    (gg1, _), (gg2, _) = run_both(emu, x, None, src, dst, n, w, ["min", "max"], ["identity"], avg)
    order = torch.sort(dst, stable=True).indices
    src_s, dst_s = src[order], dst[order]
    want = torch.zeros(n, f)
    for r in range(n):
        sl = (dst_s == r).nonzero().flatten()
        if sl.numel() == 0:
            continue
        m = x[src_s[sl]]
        for j in range(f):
            want[src_s[sl][(m[:, j] == m[:, j].min()).nonzero()[0, 0]], j] += 1.0
            want[src_s[sl][(m[:, j] == m[:, j].max()).nonzero()[0, 0]], j] += 1.0
    assert torch.equal(gg1, want)
    assert torch.equal(gg2, want)


def test_emulated_backward_bf16_rows(emu):
    n, e, f = 40, 300, 16
    src, dst, g = graph(n, e, 50, seed=4)
    x = torch.randn(n, f, generator=g).to(torch.bfloat16)
    aggrs, scalers = ["mean", "std", "sum"], ["identity", "amplification", "attenuation"]
    avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(dst, minlength=n)))
    w = torch.randn(n, 9 * f, generator=g).to(torch.bfloat16)
    (gg1, _), (gg2, _) = run_both(emu, x, None, src, dst, n, w.float(), aggrs, scalers, avg)
    want_x, _ = reference_grads(x.float(), None, src, dst, n, w.float(), aggrs, scalers, avg)
    torch.testing.assert_close(gg1, want_x, rtol=1e-3, atol=5e-4)
    torch.testing.assert_close(gg2, want_x, rtol=1e-3, atol=5e-4)


def test_emulated_backward_random_shapes(emu):
    """Both paths over random widths, tower counts, aggregator / scaler subsets, with and without row_bias, graphs with and
    without split rows: every launch geometry of launch_bwd_typed (G = 1 ... 32 lanes per row, vector and scalar rows)."""
    import random
    rnd = random.Random(0)
    for it in range(80):
        towers = rnd.choice([1, 1, 1, 2, 3, 5])
        f = rnd.choice([1, 2, 3, 4, 5, 7, 8, 12, 15, 16, 20, 33]) * towers
        n, e = rnd.randint(3, 70), rnd.randint(1, 500)
        g = torch.Generator().manual_seed(it)
        src = torch.randint(0, n, (e,), generator=g)
        dst = torch.randint(0, max(1, n - rnd.randint(0, 3)), (e,), generator=g)
        if e > 100 and rnd.random() < 0.5:
            dst[:rnd.randint(SPLIT, min(e, 200))] = rnd.randint(0, n - 1)
        x = torch.randn(n, f, generator=g)
        bias = torch.randn(n, f, generator=g) if rnd.random() < 0.5 else None
        aggrs, scalers = rnd.sample(AGGRS, rnd.randint(1, 6)), rnd.sample(SCALERS, rnd.randint(1, 5))
        avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(dst, minlength=n)))
        w = torch.randn(n, len(aggrs) * len(scalers) * f, generator=g)
        (gg1, gb1), (gg2, gb2) = run_both(emu, x, bias, src, dst, n, w, aggrs, scalers, avg, towers)
        want_x, want_b = reference_grads(x, bias, src, dst, n, w, aggrs, scalers, avg, towers)
        what = f"case {it}: F={f} towers={towers} N={n} E={e} {aggrs} {scalers} bias={bias is not None}"
        for got in (gg1, gg2):
            torch.testing.assert_close(got, want_x, rtol=2e-3, atol=2e-3, msg=lambda m: f"{what}\n{m}")
        if bias is not None:
            has_in = torch.bincount(dst, minlength=n) > 0
            for got in (gb1, gb2):
                torch.testing.assert_close(got[has_in], want_b[has_in], rtol=2e-3, atol=5e-3, msg=lambda m: f"{what}\n{m}")
                assert torch.equal(got[~has_in], torch.zeros_like(got[~has_in]))
