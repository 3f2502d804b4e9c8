"""The pull data plane on the HOST: pna_halo_pull (csrc/pna_peer.cu) executed thread by thread (tests/emu) for W "ranks" in one
process, on the plans pna_b200/dist.py builds -- the owner|row encoding, the per-rank row pitch, the 16-byte-chunk and the
narrow variants, the grid-stride loop (the emulated device has 2 "SMs").  Each rank's halo tail must equal the owners' rows,
and the rank's rows aggregated from [local ; halo] (CPU oracle) must equal the same rows of the whole graph."""
import ctypes as C
import importlib.util
import os
import shutil

import pytest
import torch

from oracle import pna_oracle as O
from pna_b200 import _lib, dist as pd, synth

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "build_emu.py"))
    build_emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_emu)
    try:
        L = C.CDLL(build_emu.build("pna_peer.cu"))
    except Exception as exc:
        pytest.skip(f"emulation library did not build: {exc}")
    L.emu_last_error.restype = C.c_char_p
    L.pna_halo_pull.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    return L


@pytest.mark.parametrize("n,e,f,world,dtype,pitch_pad", [(600, 5000, 128, 4, torch.float32, 0), (500, 3000, 256, 3, torch.float32, 0),
                                                         (400, 2500, 75, 2, torch.float32, 0), (300, 2000, 75, 3, torch.bfloat16, 0),
                                                         (300, 2000, 80, 4, torch.bfloat16, 8), (200, 1500, 520, 2, torch.float32, 0),
                                                         (64, 400, 8, 8, torch.float32, 4)])
def test_emulated_halo_pull_fills_every_ranks_halo(emu, n, e, f, world, dtype, pitch_pad):
    g = torch.Generator().manual_seed(n + f)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, int(n * 0.9), (e,), generator=g)
    x = synth.hash_features(torch.arange(n), f, dtype=dtype)
    deg = torch.bincount(dst, minlength=n)
    avg = O.avg_deg_from_histogram(torch.bincount(deg))
    bounds = pd.partition_bounds(deg, world)
    plans = []
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        mine = (dst >= lo) & (dst < hi)
        plans.append(pd.build_pull_plan(src[mine], dst[mine], bounds, r, world))
    rows = max(p.n_local + p.n_halo for p in plans)
    ld = f + pitch_pad
    bufs = [torch.full((rows, ld), -7.0, dtype=dtype) for _ in range(world)]        # every rank's [local ; halo] buffer
    for r, p in enumerate(plans):
        bufs[r][: p.n_local, :f] = x[p.lo:p.hi]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64)
    want_all = O.simple_propagate(x.float(), torch.stack([src, dst]), A4, S3, avg)
    for r, p in enumerate(plans):
        if p.n_halo:
            rc = emu.pna_halo_pull(table.data_ptr(), ld, p.enc.data_ptr(), p.shift, p.n_halo, bufs[r][p.n_local:].data_ptr(), ld, f,
                                   _lib.PNA_F32 if dtype == torch.float32 else _lib.PNA_BF16, None)
            assert rc == 0, emu.emu_last_error()
        ext = bufs[r][: p.n_local + p.n_halo]
        assert torch.equal(ext[p.n_local:, :f], x[p.halo_ids])                       # the owners' rows, bit for bit
        assert bool((ext[:, f:] == -7.0).all()) and bool((bufs[r][p.n_local + p.n_halo:] == -7.0).all())   # nothing else touched
        got = O.simple_propagate(ext[:, :f].float(), torch.stack([p.src_ext, p.dst_local]), A4, S3, avg)[: p.n_local]
        torch.testing.assert_close(got, want_all[p.lo:p.hi], rtol=1e-6, atol=1e-6)
    assert sum(p.n_remote_edges for p in plans) > 0
