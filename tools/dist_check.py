"""torchrun --nproc-per-node N tools/dist_check.py : multi-GPU parity of the peer and halo paths against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import pna_oracle as O
from pna_b200 import dist as pd

A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (n, e, f, hubdeg) in [(4000, 40000, 128, 3000), (3000, 20000, 64, 0), (5000, 30000, 256, 600)]:
    g = torch.Generator().manual_seed(n + f)
    src = torch.randint(0, n, (e,), generator=g); dst = torch.randint(0, int(n * 0.95), (e,), generator=g)
    if hubdeg:
        src = torch.cat([src, torch.randint(0, n, (hubdeg,), generator=g)]); dst = torch.cat([dst, torch.full((hubdeg,), 11)])
    x = torch.randn(n, f, generator=g)
    deg = torch.bincount(dst, minlength=n)
    bounds = pd.partition_bounds(deg, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = (dst >= lo) & (dst < hi)
    avg = O.avg_deg_from_histogram(torch.bincount(deg))
    want = O.simple_propagate(x, torch.stack([src, dst]), A4, S3, avg)[lo:hi]
    want64 = O.simple_propagate(x.double(), torch.stack([src, dst]), A4, S3, avg)[lo:hi]
    def check(got, tag):
        global ok
        got = got.cpu()
        light = deg[lo:hi] < 256
        good = torch.allclose(got[light], want[light], rtol=1e-5, atol=1e-5) and torch.allclose(got[~light].double(), want64[~light], rtol=1e-5, atol=1e-5)
        ok &= good
        print(f"rank {rank} n={n} f={f} {tag}: max err {(got - want).abs().max().item():.2e} {'ok' if good else 'MISMATCH'}", flush=True)
    # peer path
    pa = pd.PeerAggregator(src[mine].to(dev), dst[mine].to(dev), bounds, rank, world, f)
    pa.x_local.copy_(x[lo:hi].to(dev)); torch.cuda.synchronize(); pa.barrier()
    check(pa.aggregate(A4, S3, avg), "peer[" + pa._keep["how"][:24] + "]")
    # halo path, overlapped and serial
    plan = pd.build_halo_plan(src[mine].to(dev), dst[mine].to(dev), bounds, rank, world)
    for ov in (True, False):
        ha = pd.HaloAggregator(plan, f, overlap=ov)
        ha.x_local.copy_(x[lo:hi].to(dev))
        check(ha.aggregate(A4, S3, avg), f"halo overlap={ov} n_halo={plan.n_halo} interior={int(plan.interior.sum())}")
    torch.cuda.synchronize(); dist.barrier(device_ids=[local])
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("DIST ALL OK" if int(t) else "DIST FAILED")
dist.destroy_process_group()
