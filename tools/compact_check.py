"""On-GPU check + timing of the compact post path (config 2): [N,4F] aggregate + pna_linear_scaled_fwd vs the [N,12F] path.

Prints per-kernel times with the L2 flushed before every call (same rule as bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pna_b200
from pna_b200 import linear as L, synth
from pna_b200.aggregate import row_scales

dev = torch.device("cuda:0")
A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
ei, x = synth.arxiv_like()
n, f = x.shape
deg_hist = synth.degree_histogram(ei[1], n)
avg = pna_b200.avg_deg_from_histogram(deg_hist)
xd, eid = x.to(dev), ei.to(dev)
csr = pna_b200.csr_from_edge_index(eid, n)
g = torch.Generator().manual_seed(0)
w = (torch.randn(f, 12 * f, generator=g) / (12 * f) ** 0.5).to(dev)
b = torch.randn(f, generator=g).to(dev)
out12 = torch.empty((n, 12 * f), device=dev)
out4 = torch.empty((n, 4 * f), device=dev)
rs = row_scales(csr, S3, avg)

flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(128 << 20, dtype=torch.float32, device=dev)


def timed(fn, k=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(k):
        flush.zero_(); flush_rd.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


full = pna_b200.aggregate_forward(xd, csr, A4, S3, avg, out=out12)
comp = pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4)
rebuilt = torch.cat([comp * rs[:, s:s + 1] for s in range(3)], 1)
print("scaled blocks bit-exact:", bool(torch.equal(rebuilt, full)))
y12 = L.linear_tf32x3(full, w, b)
y4 = L.linear_scaled_tf32x3(comp, rs, w, b)
ref = full.double() @ w.double().t() + b.double()
print(f"max|y_scaled - fp64| {float((y4.double() - ref).abs().max()):.2e}   max|y_12f - fp64| {float((y12.double() - ref).abs().max()):.2e}"
      f"   max|y_scaled - y_12f| {float((y4 - y12).abs().max()):.2e}   max|ref| {float(ref.abs().max()):.2f}")
ok = float((y4.double() - ref).abs().max()) < 3e-5 * max(1.0, float(ref.abs().max()))

t_a12 = timed(lambda: pna_b200.aggregate_forward(xd, csr, A4, S3, avg, out=out12))
t_a4 = timed(lambda: pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4))
t_l12 = timed(lambda: L.linear_tf32x3(full, w, b))
t_l4 = timed(lambda: L.linear_scaled_tf32x3(comp, rs, w, b))
t_both12 = timed(lambda: L.linear_tf32x3(pna_b200.aggregate_forward(xd, csr, A4, S3, avg, out=out12), w, b))
t_both4 = timed(lambda: L.linear_scaled_tf32x3(pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4), rs, w, b))
print(f"aggregate [N,12F] {t_a12:.3f} ms   [N,4F] {t_a4:.3f} ms")
print(f"linear    12F in  {t_l12:.3f} ms   scaled 4F in {t_l4:.3f} ms")
print(f"both      12F     {t_both12:.3f} ms   compact {t_both4:.3f} ms   ({n} nodes, {ei.size(1)} edges, F={f})")
print("COMPACT OK" if ok and torch.equal(rebuilt, full) else "COMPACT FAILED")
