"""Per-step wall time and a profiler table of PNAConvSimple.forward_host (the bench's e2e leg), same order of operations as bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
import pna_b200
from pna_b200 import synth
import bench_common as bc
A4, S3 = bc.AGGRS, bc.SCALERS
dev = torch.device("cuda:0")
ei, x = synth.arxiv_like(n_feat=128, seed=0)
n, f = x.shape
deg_hist = synth.degree_histogram(ei[1], n)
avg = pna_b200.avg_deg_from_histogram(deg_hist)
xd, eid = x.to(dev), ei.to(dev)
csr = pna_b200.build_csr(eid[0], eid[1], n)
out = torch.empty((n, 12 * f), dtype=torch.float32, device=dev)
flush = bc.L2Flush(dev)
if "--like-bench" in sys.argv:       # the allocations / frees the bench does before its e2e leg
    ts = bc.timed_steps(lambda: pna_b200.aggregate_forward(xd, csr, A4, S3, avg, out=out), 20, 5, flush)
    par = bc.sampled_parity(out, csr.rowptr, csr.col, lambda idx: x[idx], avg, csr.split_threshold, n_rows_sample=n, max_edges=1 << 40, rows=torch.arange(n))
    print("parity", par["ok"], flush=True)
torch.manual_seed(0)
lay = pna_b200.PNAConvSimple(f, f, A4, S3, deg_hist).to(dev)
xh = x.pin_memory(); outh = torch.empty((n, f), dtype=torch.float32).pin_memory()
eihs = [ei.clone().pin_memory() for _ in range(4)]
def step(i): lay.forward_host(xh, eihs[i % 4], out=outh)
for i in range(3): step(i)
torch.cuda.synchronize()
walls = []
for i in range(20):
    t0 = time.perf_counter(); step(i); torch.cuda.synchronize(); walls.append(1e3 * (time.perf_counter() - t0))
print("per-step wall ms:", " ".join(f"{w:.2f}" for w in walls), flush=True)
print("memory: allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
st = torch.cuda.memory_stats()
print("cudaMalloc retries", st.get("num_alloc_retries"), "segments", st.get("segment.all.current"), "num_ooms", st.get("num_ooms"))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(5): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))
