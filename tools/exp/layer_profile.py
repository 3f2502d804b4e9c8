"""Kernel-level profile (torch.profiler) of the ZINC-shaped PNAConv(75,75,towers=5) forward and of PNAConvSimple at config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
import pna_b200
from pna_b200 import synth
A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
dev = torch.device("cuda:0")
ei, x, _ = synth.zinc_like(dtype=torch.float32)
n = x.size(0)
lay = pna_b200.PNAConv(75, 75, A4, S3, synth.degree_histogram(ei[1], n), towers=5, divide_input=True).to(dev)
xd, eid = x.to(dev), ei.to(dev)
csr = pna_b200.build_csr(eid[0], eid[1], n)
with torch.no_grad():
    for _ in range(5): lay(xd, eid, csr=csr)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10): lay(xd, eid, csr=csr)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
