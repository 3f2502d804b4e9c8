"""Destination-partitioned multi-GPU PNA aggregation (BASELINE.json configs[4]).

The reference has no distributed code at all ("NCCL / collective call sites: None").  Rows
(destinations) are independent, so the graph is cut into contiguous destination ranges, one per GPU / process.  A rank
owns the features and the output rows of its range and all in-edges of its rows; sources may live on other ranks.
Three ways to reach remote source rows, all behind ``pna_aggregate_fwd``:

* ``pull`` (``PullAggregator``, what ``bench.py --gpus N`` times): the rank's DE-DUPLICATED remote sources are listed once
  per graph, locally (the puller names the rows; no id exchange); per layer a device-side flag barrier
  (``pna_peer_barrier``) and ONE kernel of NVLink peer loads (``pna_halo_pull``) fill the halo tail of the rank's
  ``[local ; halo]`` buffer, and the aggregation gathers from local HBM only.  A remote row crosses NVLink once per layer
  however often it is gathered -- what a power-law graph needs.
* ``halo`` (``HaloAggregator``; the north star's wording, measured beside pull): the same rows through a pack kernel
  (``pna_gather_rows``) and ONE NCCL all-to-all-v (``torch.distributed.all_to_all_single``); rows whose sources are all
  local can be reduced while the all-to-all is in flight (masked light views).
* ``peer`` (``PeerAggregator``): ``col`` encodes ``owner << shift | row`` and the aggregation kernel gathers remote rows
  straight from the owner's HBM with the same asynchronous copies it uses for local rows -- gather and exchange are ONE
  kernel, no halo buffer at all; every remote EDGE crosses the link (peer lines are not cached in the local L2), so it
  fits graphs whose remote rows are rarely reused.

Host-side planning below is plain torch and runs on CPU tensors too (gloo), which is how tests/test_dist_cpu.py covers
it without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .aggregate import aggregate_forward
from .csr import LightView, build_csr


# ---- partitioning ------------------------------------------------------------------------------------------------
def partition_bounds(in_degree: torch.Tensor, world: int, row_cost: int = 12) -> torch.Tensor:
    """Contiguous destination ranges of equal cost (in-edges + row_cost per row), int64 [world+1]."""
    n = in_degree.numel()
    cost = torch.cumsum(in_degree.to(torch.int64) + row_cost, 0)
    total = int(cost[-1]) if n else 0
    targets = torch.arange(1, world, dtype=torch.int64) * total // world
    cuts = torch.searchsorted(cost, targets, right=False) + 1 if n else torch.zeros(world - 1, dtype=torch.int64)
    return torch.cat([torch.zeros(1, dtype=torch.int64), cuts.clamp(max=n), torch.tensor([n], dtype=torch.int64)])


def owner_of(ids: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
    return torch.bucketize(ids, bounds[1:].to(ids.device), right=True)


def peer_shift_for(bounds: torch.Tensor) -> int:
    biggest = int((bounds[1:] - bounds[:-1]).max())
    shift = max(1, (max(biggest, 1) - 1).bit_length())
    world = bounds.numel() - 1
    if shift > 30 or (world << shift) >= 2 ** 31:
        raise ValueError("partition too large for the 32-bit owner|row encoding")
    return shift


def encode_peer_sources(src_global: torch.Tensor, bounds: torch.Tensor, shift: int) -> torch.Tensor:
    """Global source id -> owner << shift | row-on-owner."""
    b = bounds.to(src_global.device)
    own = owner_of(src_global, bounds)
    return (own << shift) | (src_global - b[own])


@dataclass
class HaloPlan:
    """What one rank needs to run its rows with a [local ; halo] source buffer."""
    rank: int
    world: int
    lo: int
    hi: int
    n_local: int
    n_halo: int
    src_ext: torch.Tensor          # int64 [E_r] sources remapped to [0, n_local + n_halo)
    dst_local: torch.Tensor        # int64 [E_r]
    halo_ids: torch.Tensor         # int64 [n_halo] global ids, sorted (hence grouped by owner)
    recv_splits: List[int]         # rows received from each rank per exchange
    send_splits: List[int]         # rows sent to each rank per exchange
    send_idx: torch.Tensor         # int32 [sum(send_splits)] local rows to send, grouped by destination rank
    interior: torch.Tensor         # bool [n_local]: every source of the row is local


def build_halo_plan(src_global: torch.Tensor, dst_global: torch.Tensor, bounds: torch.Tensor, rank: int, world: int,
                    group=None) -> HaloPlan:
    """src_global -> dst_global are the in-edges of this rank's rows (dst in [bounds[rank], bounds[rank+1]))."""
    dev = src_global.device
    b = bounds.to(dev)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    n_local = hi - lo
    if src_global.numel() and (int(dst_global.min()) < lo or int(dst_global.max()) >= hi):
        raise ValueError("build_halo_plan: an edge's destination is outside this rank's range")
    dst_local = dst_global - lo
    remote = (src_global < lo) | (src_global >= hi)
    halo_ids = torch.unique(src_global[remote])                                   # sorted
    n_halo = int(halo_ids.numel())
    pos = torch.searchsorted(halo_ids, src_global.clamp(min=0)) if n_halo else torch.zeros_like(src_global)
    src_ext = torch.where(remote, n_local + pos, src_global - lo)
    own = owner_of(halo_ids, bounds)
    recv_counts = torch.bincount(own, minlength=world)
    # tell every owner which of its rows we need: counts first, then the (owner-local) ids
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts, group=group)
    want = (halo_ids - b[own]).to(torch.int64)
    send_idx = torch.empty(int(send_counts.sum()), dtype=torch.int64, device=dev)
    dist.all_to_all_single(send_idx, want, output_split_sizes=send_counts.tolist(), input_split_sizes=recv_counts.tolist(),
                           group=group)
    has_remote = torch.zeros(n_local, dtype=torch.bool, device=dev)
    if src_global.numel():
        has_remote.index_put_((dst_local[remote],), torch.ones(1, dtype=torch.bool, device=dev).expand(int(remote.sum())))
    return HaloPlan(rank, world, lo, hi, n_local, n_halo, src_ext, dst_local, halo_ids, recv_counts.tolist(),
                    send_counts.tolist(), send_idx.to(torch.int32), ~has_remote)


# ---- device side ---------------------------------------------------------------------------------------------------
def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[i] = src[idx[i]] through the C ABI (pna_gather_rows): packs the all-to-all send buffer."""
    if idx.numel() == 0:
        return out
    dt = {torch.float32: _lib.PNA_F32, torch.bfloat16: _lib.PNA_BF16}[src.dtype]
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().pna_gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), idx.numel(), out.data_ptr(),
                                               out.stride(0), src.size(1), dt, torch.cuda.current_stream(src.device).cuda_stream))
    return out


class HaloAggregator:
    """[local ; halo] path: pack -> one NCCL all-to-all-v -> aggregation, interior rows overlapped with the exchange."""

    def __init__(self, plan: HaloPlan, n_feat: int, dtype=torch.float32, group=None, overlap: bool = True):
        dev = plan.src_ext.device
        self.plan, self.group, self.overlap = plan, group, overlap
        self.csr = build_csr(plan.src_ext, plan.dst_local, plan.n_local, n_src=plan.n_local + plan.n_halo)
        self.x_ext = torch.zeros((plan.n_local + plan.n_halo, n_feat), dtype=dtype, device=dev)
        self.send_buf = torch.empty((int(plan.send_idx.numel()), n_feat), dtype=dtype, device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.view_interior: Optional[LightView] = None
        self.view_boundary: Optional[LightView] = None
        if overlap:
            self.view_interior = self.csr.masked_view(plan.interior)
            self.view_boundary = self.csr.masked_view(~plan.interior)

    @property
    def x_local(self) -> torch.Tensor:
        """The rank's own feature rows: produce the layer input in place here (head of the [local ; halo] buffer)."""
        return self.x_ext[: self.plan.n_local]

    def exchange(self) -> None:
        p = self.plan
        gather_rows(self.x_local, p.send_idx, self.send_buf)
        dist.all_to_all_single(self.x_ext[p.n_local:], self.send_buf, output_split_sizes=p.recv_splits,
                               input_split_sizes=p.send_splits, group=self.group)

    def aggregate(self, aggregators, scalers, avg_deg, out: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        main = torch.cuda.current_stream(self.x_ext.device)
        if not self.overlap:
            self.exchange()
            return aggregate_forward(self.x_ext, self.csr, aggregators, scalers, avg_deg, out=out, **kw)
        self.comm_stream.wait_stream(main)                     # x_local is produced on the main stream
        with torch.cuda.stream(self.comm_stream):
            self.exchange()
        out = aggregate_forward(self.x_ext, self.csr, aggregators, scalers, avg_deg, out=out, view=self.view_interior,
                                skip_hubs=True, **kw)          # rows with local sources only: no dependence on the halo
        main.wait_stream(self.comm_stream)
        return aggregate_forward(self.x_ext, self.csr, aggregators, scalers, avg_deg, out=out, view=self.view_boundary, **kw)


@dataclass
class PullPlan:
    """What one rank needs for the pull plane: computed locally from the rank's own in-edges, no id exchange at all
    (the puller names the rows; the owners do nothing)."""
    rank: int
    world: int
    lo: int
    hi: int
    n_local: int
    n_halo: int
    shift: int
    src_ext: torch.Tensor          # int64 [E_r] sources remapped to [0, n_local + n_halo)
    dst_local: torch.Tensor        # int64 [E_r]
    halo_ids: torch.Tensor         # int64 [n_halo] global ids of the de-duplicated remote sources, sorted (grouped by owner)
    enc: torch.Tensor              # int32 [n_halo] owner << shift | row-on-owner
    n_remote_edges: int


def build_pull_plan(src_global: torch.Tensor, dst_global: torch.Tensor, bounds: torch.Tensor, rank: int, world: int) -> PullPlan:
    dev = src_global.device
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    n_local = hi - lo
    if src_global.numel() and (int(dst_global.min()) < lo or int(dst_global.max()) >= hi):
        raise ValueError("build_pull_plan: an edge's destination is outside this rank's range")
    remote = (src_global < lo) | (src_global >= hi)
    halo_ids = torch.unique(src_global[remote])
    n_halo = int(halo_ids.numel())
    pos = torch.searchsorted(halo_ids, src_global) if n_halo else torch.zeros_like(src_global)
    src_ext = torch.where(remote, n_local + pos, src_global - lo)
    shift = peer_shift_for(bounds)
    enc = encode_peer_sources(halo_ids, bounds, shift).to(torch.int32) if n_halo else torch.zeros(0, dtype=torch.int32, device=dev)
    return PullPlan(rank, world, lo, hi, n_local, n_halo, shift, src_ext, dst_global - lo, halo_ids, enc, int(remote.sum()))


class PullAggregator:
    """[local ; halo] source buffer whose halo tail is filled by ONE kernel of peer loads (``pna_halo_pull``): the
    all-to-all of the north star without a collective -- no id exchange when the graph is planned, no pack kernel, no
    send buffer, no NCCL call per layer; a remote row crosses NVLink once per layer however often it is gathered.

    Protocol (one layer): every rank writes its rows into ``x_local`` -> ``exchange()`` = device-side barrier
    (``pna_peer_barrier``: one flag store per peer, spin on the own flags) + the pull -> aggregation from the local
    buffer.  The feature buffer is DOUBLE-BUFFERED (``flip()`` between layers / steps): a rank may already be writing
    layer l+1's rows while slower peers still pull layer l's, and the barrier of layer l+1 separates layer l's pulls
    from the writes of layer l+2 into the same buffer.
    """

    def __init__(self, plan: PullPlan, n_feat: int, dtype=torch.float32, group=None, buffers: int = 2, _alloc=None):
        dev = plan.src_ext.device
        self.plan, self.group, self.n_feat, self.dtype = plan, group, n_feat, dtype
        self.csr = build_csr(plan.src_ext, plan.dst_local, plan.n_local, n_src=plan.n_local + max(plan.n_halo, 0))
        rows = torch.tensor([plan.n_local + plan.n_halo], dtype=torch.int64, device=dev)
        if _alloc is None and plan.world > 1:
            dist.all_reduce(rows, op=dist.ReduceOp.MAX, group=group)
        alloc = _alloc or (lambda shape, dt: _symmetric_tensor(shape, dt, dev, plan.rank, plan.world, group))
        self._bufs, self._tables, self._keep = [], [], []
        for _ in range(buffers):
            t, ptrs, keep = alloc((int(rows), n_feat), dtype)
            self._bufs.append(t)
            self._tables.append(torch.tensor(ptrs, dtype=torch.int64, device=dev))
            self._keep.append(keep)
        flags, fptrs, keep = alloc((max(plan.world, 1),), torch.int64)
        flags.zero_()
        self._flags, self._flag_table = flags, torch.tensor(fptrs, dtype=torch.int64, device=dev)
        self._keep.append(keep)
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._epoch, self._cur = 0, 0
        self.use_barrier = plan.world > 1 and _alloc is None
        if self.use_barrier:   # flags are zero everywhere before the first flag store can arrive
            torch.cuda.synchronize(dev)
            dist.barrier(group=group, device_ids=[dev.index])

    @property
    def x_ext(self) -> torch.Tensor:
        return self._bufs[self._cur][: self.plan.n_local + self.plan.n_halo]

    @property
    def x_local(self) -> torch.Tensor:
        """This rank's own feature rows: produce the layer input in place here (head of the [local ; halo] buffer)."""
        return self._bufs[self._cur][: self.plan.n_local]

    def flip(self) -> None:
        self._cur = (self._cur + 1) % len(self._bufs)

    def barrier(self) -> None:
        self._epoch += 1
        dev = self._flags.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pna_peer_barrier(self._flag_table.data_ptr(), self.plan.rank, self.plan.world, self._epoch, 0,
                                                   self._status.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))

    def exchange(self) -> None:
        p = self.plan
        if self.use_barrier:
            self.barrier()
        if p.n_halo == 0:
            return
        buf = self._bufs[self._cur]
        dt = {torch.float32: _lib.PNA_F32, torch.bfloat16: _lib.PNA_BF16}[self.dtype]
        dev = buf.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pna_halo_pull(self._tables[self._cur].data_ptr(), buf.stride(0), p.enc.data_ptr(), p.shift, p.n_halo,
                                                buf[p.n_local:].data_ptr(), buf.stride(0), self.n_feat, dt,
                                                torch.cuda.current_stream(dev).cuda_stream))

    def aggregate(self, aggregators, scalers, avg_deg, out: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        self.exchange()
        return aggregate_forward(self.x_ext, self.csr, aggregators, scalers, avg_deg, out=out, **kw)

    def check(self) -> None:
        """Host-side check (synchronises): did every barrier see all peers arrive?"""
        if int(self._status.item()) != 0:
            raise RuntimeError("pna_peer_barrier timed out: a peer rank did not reach the barrier")


def _symmetric_tensor(shape, dtype, dev, rank: int, world: int, group):
    """A tensor of `shape` on every rank, each mapped into every process; (local tensor, per-rank pointers, keepalive)."""
    numel = 1
    for d in shape:
        numel *= int(d)
    rows = _symmetric_rows(1, max(numel, 1), dtype, dev, rank, world, group)
    return rows[0].view(-1)[:numel].view(*shape), rows[1], rows[2]


class PeerAggregator:
    """Gather fused with the exchange: remote rows are read over NVLink inside the aggregation kernel.

    Protocol: every rank writes its rows into ``x_local`` -> ``barrier()`` -> ``aggregate()``.  The barrier orders "all ranks
    have written" before the gathers; NOTHING orders the end of the peers' gathers before this rank's next write into
    ``x_local`` -- a rank that finishes ``aggregate()`` early must not overwrite its rows while slower peers may still be
    reading them.  Either call ``barrier()`` again after ``aggregate()`` before rewriting ``x_local`` (what a multi-layer net
    using ONE buffer has to do), or alternate between two aggregators / buffers per layer as ``PullAggregator`` does
    (``flip()``), where the next layer's barrier separates this layer's reads from the writes two layers later."""

    def __init__(self, src_global: torch.Tensor, dst_global: torch.Tensor, bounds: torch.Tensor, rank: int, world: int,
                 n_feat: int, dtype=torch.float32, group=None):
        dev = src_global.device
        self.rank, self.world, self.group = rank, world, group
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        self.n_local = hi - lo
        self.shift = peer_shift_for(bounds)
        enc = encode_peer_sources(src_global, bounds, self.shift)
        self.csr = build_csr(enc, dst_global - lo, self.n_local, n_src=world << self.shift)
        rows_max = int((bounds[1:] - bounds[:-1]).max())
        self.x_local, self.peer_ptrs, self._keep = _symmetric_rows(rows_max, n_feat, dtype, dev, rank, world, group)
        self.x_local = self.x_local[: self.n_local]
        self.ptr_table = torch.tensor(self.peer_ptrs, dtype=torch.int64, device=dev)

    def barrier(self) -> None:
        """All ranks have finished writing their x rows (device-side, on the current stream)."""
        h = self._keep.get("handle")
        if h is not None and hasattr(h, "barrier"):
            h.barrier()
        else:
            dist.barrier(group=self.group, device_ids=[self.x_local.device.index])

    def aggregate(self, aggregators, scalers, avg_deg, out: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        return aggregate_forward(self.x_local, self.csr, aggregators, scalers, avg_deg, out=out,
                                 peer=(self.ptr_table, self.shift), **kw)


def _symmetric_rows(rows: int, n_feat: int, dtype, dev, rank: int, world: int, group):
    """A [rows, n_feat] buffer on every rank, each mapped into every process; returns (local tensor, pointers, keepalive)."""
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty((rows, n_feat), dtype=dtype, device=dev)
        h = symm.rendezvous(t, group=group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in h.buffer_ptrs]
        return t, ptrs, {"handle": h, "tensor": t, "how": "torch symmetric memory"}
    except Exception as exc:  # CUDA IPC fallback: share the caching-allocator block, open it on every peer
        t = torch.empty((rows, n_feat), dtype=dtype, device=dev)
        meta = t.untyped_storage()._share_cuda_()
        metas = [None] * world
        dist.all_gather_object(metas, meta, group=group)
        opened, ptrs = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(t.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*metas[r])
            peer = torch.empty(0, dtype=dtype, device=st.device).set_(st, 0, (rows, n_feat), (n_feat, 1))
            opened.append((st, peer))
            ptrs.append(peer.data_ptr())
        return t, ptrs, {"opened": opened, "tensor": t, "how": f"CUDA IPC (symmetric memory unavailable: {exc})"}


# ---- synthetic destination-partitioned workload (tools/dist_check.py; the bench lines come from bench_multi.py) ------
def rank_graph(rank: int, world: int, n_local: int, e_local: int, n_feat: int, p_remote: float, seed: int = 0,
               skew: float = 3.0, dtype=torch.float32):
    """In-edges of rank's rows in a graph of world * n_local nodes: destinations skewed like synth.arxiv_like inside
    the rank's range; a source is drawn from the whole graph with probability p_remote, else from the rank's range."""
    g = torch.Generator().manual_seed(seed * 1000 + rank)
    lo = rank * n_local
    perm = torch.randperm(n_local, generator=g)
    u = torch.rand(e_local, generator=g, dtype=torch.float64)
    dst = lo + perm[(n_local * u.pow(skew)).long().clamp_(max=n_local - 1)]
    local_src = lo + torch.randint(0, n_local, (e_local,), generator=g)
    any_src = torch.randint(0, world * n_local, (e_local,), generator=g)
    src = torch.where(torch.rand(e_local, generator=g) < p_remote, any_src, local_src)
    x = torch.randn(n_local, n_feat, generator=g).to(dtype)
    return src, dst, x
