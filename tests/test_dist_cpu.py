"""Host-side logic of the destination-partitioned multi-GPU path on CPU: world_size-2 gloo processes build the halo
plan / the peer encoding and the oracle is evaluated through them; the result must equal the oracle on the whole graph."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
A4 = ["mean", "max", "min", "std"]
S3 = ["identity", "amplification", "attenuation"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _graph(n=240, e=2500, f=6, seed=5):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, int(n * 0.9), (e,), generator=g)
    dst[: e // 5] = 7                       # a heavy destination so the cost-balanced cut is not the midpoint
    x = torch.randn(n, f, generator=g)
    return src, dst, x


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pna_oracle as O
        from pna_b200 import dist as pd
        src, dst, x = _graph()
        n = x.size(0)
        deg = torch.bincount(dst, minlength=n)
        bounds = pd.partition_bounds(deg, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = (dst >= lo) & (dst < hi)
        avg = O.avg_deg_from_histogram(torch.bincount(deg))
        want = O.simple_propagate(x, torch.stack([src, dst]), A4, S3, avg)[lo:hi]

        # --- halo path: plan -> exchange rows with gloo -> oracle on [local ; halo]
        plan = pd.build_halo_plan(src[mine], dst[mine], bounds, rank, world)
        x_local = x[lo:hi]
        send = x_local[plan.send_idx.long()]
        recv = torch.empty((plan.n_halo, x.size(1)))
        dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_splits, input_split_sizes=plan.send_splits)
        assert torch.equal(recv, x[plan.halo_ids])                        # the right rows arrive in the right order
        x_ext = torch.cat([x_local, recv])
        # the oracle scatters into n_local + n_halo rows; only the first n_local are this rank's destinations
        got = O.simple_propagate(x_ext, torch.stack([plan.src_ext, plan.dst_local]), A4, S3, avg)[: plan.n_local]
        assert torch.equal(got, want)
        # interior rows really have no remote source
        rem = (src[mine] < lo) | (src[mine] >= hi)
        touched = torch.zeros(plan.n_local, dtype=torch.bool)
        touched[plan.dst_local[rem]] = True
        assert torch.equal(plan.interior, ~touched)

        # --- peer path: owner << shift | row decodes back to the global source
        shift = pd.peer_shift_for(bounds)
        enc = pd.encode_peer_sources(src[mine], bounds, shift)
        own, row = enc >> shift, enc & ((1 << shift) - 1)
        assert torch.equal(bounds[own] + row, src[mine]) and int(own.max()) < world
        q.put((rank, "ok", int(plan.n_halo), bounds.tolist()))
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, []))
    finally:
        dist.destroy_process_group()


def test_halo_plan_and_peer_encoding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[1]
    assert all(r[2] > 0 for r in res)                 # both ranks need remote rows
    b = res[0][3]
    assert b[0] == 0 and b[-1] == 240 and b[1] != 120  # cost-balanced, not the midpoint


def test_partition_bounds_balance():
    from pna_b200 import dist as pd
    deg = torch.cat([torch.full((100,), 50), torch.zeros(900, dtype=torch.long)])
    b = pd.partition_bounds(deg, 4)
    cost = torch.cumsum(deg + 12, 0)
    parts = [int(cost[b[i + 1] - 1] - (cost[b[i] - 1] if b[i] > 0 else 0)) for i in range(4)]
    assert max(parts) - min(parts) <= 2 * 62 and b[0] == 0 and b[-1] == 1000
