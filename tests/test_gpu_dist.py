"""The peer-memory data plane on ONE GPU: W "ranks" live in one process, every rank's feature buffer is an ordinary device
tensor and the pointer table names them all -- pna_halo_pull and the peer gather only see pointers, so the code path is
the one the multi-GPU run takes over NVLink (bench_multi.py asserts parity of the real thing in the same run)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

A4 = ["mean", "max", "min", "std"]
S3 = ["identity", "amplification", "attenuation"]


def dev():
    return torch.device("cuda:0")


def _graph(n, e, hub, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, int(n * 0.93), (e,), generator=g)
    if hub:
        src = torch.cat([src, torch.randint(0, n, (hub,), generator=g)])
        dst = torch.cat([dst, torch.full((hub,), n // 3)])
        p = torch.randperm(src.numel(), generator=g)
        src, dst = src[p], dst[p]
    return src, dst


@pytest.mark.parametrize("n,e,hub,f,world,dtype", [(3000, 30000, 2500, 128, 4, torch.float32), (1500, 9000, 0, 256, 3, torch.float32),
                                                   (2000, 12000, 700, 64, 2, torch.float32), (1200, 8000, 0, 80, 4, torch.bfloat16),
                                                   (900, 5000, 300, 75, 3, torch.float32)])
def test_pull_plane_equals_single_gpu_bit_for_bit(n, e, hub, f, world, dtype):
    import pna_b200
    from pna_b200 import dist as pd, synth
    src, dst = _graph(n, e, hub, seed=n + f)
    x = synth.hash_features(torch.arange(n), f, dtype=dtype)
    deg = torch.bincount(dst, minlength=n)
    avg = pna_b200.avg_deg_from_histogram(torch.bincount(deg))
    bounds = pd.partition_bounds(deg, world)
    csr = pna_b200.build_csr(src.to(dev()), dst.to(dev()), n)
    want = pna_b200.aggregate_forward(x.to(dev()), csr, A4, S3, avg)

    plans = []
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        mine = (dst >= lo) & (dst < hi)
        plans.append(pd.build_pull_plan(src[mine].to(dev()), dst[mine].to(dev()), bounds, r, world))
    rows = max(p.n_local + p.n_halo for p in plans)
    bufs = [torch.zeros((rows, f), dtype=dtype, device=dev()) for _ in range(world)]
    flags = [torch.zeros(world, dtype=torch.int64, device=dev()) for _ in range(world)]

    def alloc_for(r):
        calls = {"i": 0}

        def alloc(shape, dt):
            calls["i"] += 1
            if len(shape) == 2:
                return bufs[r], [b.data_ptr() for b in bufs], None
            return flags[r], [fl.data_ptr() for fl in flags], None
        return alloc
    aggs = [pd.PullAggregator(plans[r], f, dtype=dtype, buffers=1, _alloc=alloc_for(r)) for r in range(world)]
    for r in range(world):
        aggs[r].x_local.copy_(x[int(bounds[r]):int(bounds[r + 1])])
    for r in range(world):
        got = aggs[r].aggregate(A4, S3, avg)
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        # same slot order, same chunking of the split rows -> the same bits as the single-GPU call
        assert torch.equal(got, want[lo:hi]), f"rank {r}"
        assert torch.equal(aggs[r].x_ext[plans[r].n_local:], x.to(dev())[plans[r].halo_ids])


def test_peer_barrier_single_rank_and_timeout():
    """world = 1 passes immediately; a missing peer trips the timeout and sets the status word instead of hanging."""
    import ctypes as C
    from pna_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    flags = torch.zeros(2, dtype=torch.int64, device=dev())
    table = torch.tensor([flags.data_ptr(), flags.data_ptr()], dtype=torch.int64, device=dev())
    status = torch.zeros(1, dtype=torch.int32, device=dev())
    _lib.check(L.pna_peer_barrier(table.data_ptr(), 0, 1, 1, 0, status.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and int(flags[0].item()) == 1
    # two ranks, the second never arrives: flags[1] stays 0 -> 5 ms timeout
    flags.zero_()
    _lib.check(L.pna_peer_barrier(table.data_ptr(), 0, 2, 1, 5_000_000, status.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(status.item()) == 1
