"""Round-2 experiment (the cheap intermediate step): does the layer get faster when the
aggregation of row block b+1 runs beside the tensor-core GEMM of row block b?

The two kernels are bound by different units (gather latency vs tensor pipe).  Uses only validated public pieces:
masked light views (one per contiguous row block, built once per graph), `aggregate_forward(view=..., skip_hubs=True)`
on stream 1, `linear_scaled_tf32x3` of the finished block on stream 2.  Prints serial vs overlapped time and checks the
results are identical.  Not run by the tests.

    python tools/probes/overlap_probe.py [n_blocks]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pna_b200
from pna_b200 import linear as L, synth
from pna_b200.aggregate import row_scales

dev = torch.device("cuda:0")
A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ei, x = synth.arxiv_like()
n, f = x.shape
avg = pna_b200.avg_deg_from_histogram(synth.degree_histogram(ei[1], n))
xd = x.to(dev)
csr = pna_b200.csr_from_edge_index(ei.to(dev), n)
g = torch.Generator().manual_seed(0)
w = (torch.randn(f, 12 * f, generator=g) / (12 * f) ** 0.5).to(dev)
b = torch.randn(f, generator=g).to(dev)
rs = row_scales(csr, S3, avg)
out4 = torch.empty((n, 4 * f), device=dev)
y_serial = torch.empty((n, f), device=dev)
y_blocks = torch.empty((n, f), device=dev)

step = (n + n_blocks - 1) // n_blocks
step = (step + 127) // 128 * 128                      # GEMM tiles are 128 rows
bounds = [(r0, min(n, r0 + step)) for r0 in range(0, n, step)]
views = []
for r0, r1 in bounds:
    m = torch.zeros(n, dtype=torch.uint8, device=dev)
    m[r0:r1] = 1
    views.append(csr.masked_view(m))
s_agg, s_gemm = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def serial():
    a = pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4)
    y_serial.copy_(L.linear_scaled_tf32x3(a, rs, w, b))


def overlapped():
    main = torch.cuda.current_stream(dev)
    s_agg.wait_stream(main); s_gemm.wait_stream(main)
    with torch.cuda.stream(s_agg):
        pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4, skip_light=True)      # split rows first
    evs = []
    for (r0, r1), v in zip(bounds, views):
        with torch.cuda.stream(s_agg):
            pna_b200.aggregate_forward(xd, csr, A4, ["identity"], avg, out=out4, view=v, skip_hubs=True)
            ev = torch.cuda.Event(); ev.record(s_agg); evs.append(ev)
        with torch.cuda.stream(s_gemm):
            s_gemm.wait_event(ev)
            y_blocks[r0:r1].copy_(L.linear_scaled_tf32x3(out4[r0:r1], rs[r0:r1], w, b))
    main.wait_stream(s_agg); main.wait_stream(s_gemm)


def timed(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / k


serial(); overlapped(); torch.cuda.synchronize()
print("identical results:", bool(torch.equal(y_serial, y_blocks)), " max diff", float((y_serial - y_blocks).abs().max()))
print(f"serial {timed(serial):.3f} ms    {len(bounds)} row blocks, two streams {timed(overlapped):.3f} ms")
