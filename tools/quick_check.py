"""Quick on-GPU sanity check of the aggregation against the oracle on a few shapes (used before the full suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pna_b200
from oracle import pna_oracle as O
from pna_b200 import synth

A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
dev = torch.device("cuda:0")
bad = 0
for (n, e, f, hub) in [(50, 300, 128, 0), (1000, 9000, 128, 0), (3000, 30000, 128, 2000), (700, 5000, 256, 0), (500, 4000, 96, 600),
                       (400, 3000, 512, 0), (2000, 1000, 128, 0), (17, 0, 128, 0)]:
    g = torch.Generator().manual_seed(n + f)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, max(1, int(n * 0.9)), (e,), generator=g)])
    if hub:
        ei = torch.cat([ei, torch.stack([torch.randint(0, n, (hub,), generator=g), torch.full((hub,), n - 1)])], 1)
    x = torch.randn(n, f, generator=g)
    avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(ei[1], minlength=n))) if ei.numel() else {"log": 1.0, "lin": 1.0}
    want = O.simple_propagate(x, ei, A4, S3, avg)
    csr = pna_b200.build_csr(ei[0].to(dev), ei[1].to(dev), n)
    t0 = time.time()
    got = pna_b200.aggregate_forward(x.to(dev), csr, A4, S3, avg)
    torch.cuda.synchronize()
    err = (got.cpu() - want).abs().max().item() if n else 0.0
    ok = torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)
    bad += not ok
    print(f"n={n} e={e} f={f} hub={hub}: max err {err:.2e} {'ok' if ok else 'MISMATCH'} ({time.time()-t0:.3f}s)", flush=True)
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
