#!/usr/bin/env python
"""bench.py -- aggregated edges/s of the PNA layer forward on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over the whole graph: CSR (resident, built once) -> [N, 12*F] aggregation
(mean/max/min/std x identity/amplification/attenuation) -- the kernels of libpna_sm100.so and nothing else.
  value        edges/s of that step, inputs resident in HBM, CUDA events around each step, L2 flushed between steps
  e2e          edges/s of PNAConvSimple.forward(x, edge_index) called with pinned HOST tensors: H2D of x and
               edge_index, CSR build, aggregation, post-MLP, D2H of the layer output, all inside the timed region
  roofline     B_min / step time against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline the reference's PyTorch CPU op sequence (oracle/pna_oracle.py, a port: torch_geometric/torch_scatter
               are not installable) timed on the host cores of this box on the same graph
N = 1: BASELINE.json configs[1] (ogbn-arxiv-shaped, 169 343 nodes / 1 166 243 edges, F = 128, fp32).
N > 1: weak scaling -- a graph N times larger, destination-partitioned, one halo all-to-all per step (pna_b200/dist.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

AGGRS = ["mean", "max", "min", "std"]
SCALERS = ["identity", "amplification", "attenuation"]
METRIC = "aggregated edges/sec (PNA layer fwd)"
UNIT = "edges/s"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(world: int, rank: int):
    """config 2 for this rank.  N = 1: the graph itself.  N > 1: see pna_b200/dist.py (weak scaling)."""
    from pna_b200 import synth
    ei, x = synth.arxiv_like(n_feat=128, seed=0)
    return ei, x


def cpu_reference_layer(ei, x, deg_hist, steps: int, warmup: int):
    """The reference's CPU path for PNAConvSimple.forward (port: oracle/pna_oracle.py) on the host cores of this box.

    torch's CPU scatter/index kernels do not scale to 100+ threads (oversubscription makes them slower), so a few
    thread counts are tried in the warm-up and the FASTEST one is timed and reported -- the baseline gets every
    advantage the hardware offers."""
    from oracle import pna_oracle as O
    f = x.size(1)
    torch.manual_seed(0)
    lay = O.PNAConvSimpleOracle(f, f, AGGRS, SCALERS, deg_hist)
    ncpu = os.cpu_count() or 1
    candidates = sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True)
    best_t, best_n = None, ncpu
    with torch.no_grad():
        for n in candidates:
            torch.set_num_threads(n)
            lay(x, ei)                                  # warm this setting
            t0 = time.perf_counter()
            lay(x, ei)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best_t, best_n = dt, n
        torch.set_num_threads(best_n)
        times = []
        for i in range(max(0, warmup - 1) + steps):
            t0 = time.perf_counter()
            lay(x, ei)
            dt = time.perf_counter() - t0
            if i >= max(0, warmup - 1):
                times.append(dt)
    return times, best_n, {"threads_tried": candidates, "host_cpus": ncpu}


def run_reference(args):
    """--impl reference: the reference arm (CPU).  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pna_b200 import synth
    ei, x = make_workload(1, 0)
    n, e = x.size(0), ei.size(1)
    deg = synth.degree_histogram(ei[1], n)
    times, threads, info = cpu_reference_layer(ei, x, deg, args.steps, max(1, min(args.warmup, 2)))
    total = sum(times)
    v = e * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ogbn-arxiv-shaped CSR (configs[1])", "n_nodes": n, "n_edges": e, "n_feat": x.size(1),
                   "layer": "PNAConvSimple(128,128) forward: aggregate + post-MLP"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", **info,
                         "sample": f"full config-2 graph, {len(times)} forward passes of the reference op sequence "
                                   "(index_select, 6x scatter_add, amin, amax, degree, 3 scalers, cats, Linear) in torch CPU"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist
    import pna_b200
    from pna_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from pna_b200 import dist as pdist
        return pdist.bench_multi_gpu(args, METRIC, UNIT, AGGRS, SCALERS, measured_peaks, ClockSampler)

    ei, x = make_workload(1, 0)
    n, e, f = x.size(0), ei.size(1), x.size(1)
    deg_hist = synth.degree_histogram(ei[1], n)
    avg_deg = pna_b200.avg_deg_from_histogram(deg_hist)
    xd, eid = x.to(dev), ei.to(dev)

    # CSR: once per graph (cached by the layers); timed separately
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csr = pna_b200.build_csr(eid[0], eid[1], n)
    torch.cuda.synchronize()
    csr_ms_first = 1e3 * (time.perf_counter() - t0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        pna_b200.build_csr(eid[0], eid[1], n)
    ev[1].record()
    torch.cuda.synchronize()
    csr_ms = ev[0].elapsed_time(ev[1]) / 5

    out = torch.empty((n, 12 * f), dtype=torch.float32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    flush_rd = torch.zeros(128 << 20, dtype=torch.float32, device=dev)   # 512 MiB, read after the write (see l2_flush)

    def l2_flush():
        """Evict everything of the previous iteration: write 512 MiB, then READ another 512 MiB so the L2 is left full of
        CLEAN lines -- a write-only flush leaves ~126 MB of dirty lines whose write-back would be charged to the timed step."""
        flush.zero_()
        flush_rd.sum()

    def step(**kw):
        pna_b200.aggregate_forward(xd, csr, AGGRS, SCALERS, avg_deg, out=out, **kw)

    def timed(k, warm, **kw):
        for _ in range(warm):
            l2_flush(); step(**kw)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        torch.cuda.synchronize()
        for i in range(k):
            l2_flush()                         # evict x / CSR / out lines of the previous iteration from L2
            starts[i].record()
            step(**kw)
            ends[i].record()
        torch.cuda.synchronize()
        return [s.elapsed_time(t) for s, t in zip(starts, ends)]

    # clocks / throttle reasons are sampled (20 Hz) while the GPU runs this workload: the timed steps themselves last
    # only ~15 ms, so the sampler brackets them with extra untimed passes of the same kernels to collect enough samples
    with ClockSampler(local) as clk:
        time.sleep(0.06)
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
        per_step = timed(args.steps, args.warmup)
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
    clocks = clk.summary()
    t_ms = sum(per_step) / len(per_step)
    value = e / (t_ms * 1e-3)
    from pna_b200.aggregate import fold_finalize_enabled
    # k_rows_stream (+ k_hub_finalize when rows were split and the finalize is not folded into the stream kernel)
    launches_per_step = 1 + (1 if (csr.n_hubs and not fold_finalize_enabled()) else 0)
    # (the e2e / layer_fwd legs additionally launch k_split_weight + k_linear_3xtf32 and the CSR-build kernels)

    bytes_ = synth.algorithmic_bytes(n, e, f, 4, 12 * f)
    peak, peak_src = measured_peaks()
    achieved = bytes_["b_min"] / (t_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_step")
        except Exception:
            traffic = None

    # e2e: the public layer call with HOST buffers (pinned), copies inside the timed region
    torch.manual_seed(0)
    lay = pna_b200.PNAConvSimple(f, f, AGGRS, SCALERS, deg_hist).to(dev)
    xh, eih = x.pin_memory(), ei.pin_memory()
    outh = torch.empty((n, f), dtype=torch.float32).pin_memory()

    def e2e_step():
        # the public host-buffer call: pinned x / edge_index in, pinned result out; a new edge_index object every step,
        # so the CSR is rebuilt inside the call (nothing is cached across steps)
        lay.forward_host(xh, eih_steps[e2e_step.i % len(eih_steps)], out=outh)
        e2e_step.i += 1
    e2e_step.i = 0
    eih_steps = [ei.clone().pin_memory() for _ in range(4)]    # distinct host tensors: the device copy is always fresh

    # PCIe health of this box (context for e2e: the layer call moves 192 MB per step over PCIe)
    def copy_rate(fn, nbytes):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        return nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
    xdev_tmp = torch.empty_like(xd)
    h2d_gbs = copy_rate(lambda: xdev_tmp.copy_(xh, non_blocking=True), xh.numel() * 4)
    ydev_tmp = torch.empty((n, f), dtype=torch.float32, device=dev)
    d2h_gbs = copy_rate(lambda: outh.copy_(ydev_tmp, non_blocking=True), n * f * 4)
    del xdev_tmp, ydev_tmp

    k2 = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s2.record()
    for _ in range(k2):
        e2e_step()
    e2.record()
    torch.cuda.synchronize()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0) / k2
    e2e_ms = max(s2.elapsed_time(e2) / k2, e2e_wall_ms)      # host-side launch/sync time counts too
    def layer_ms():
        with torch.no_grad():
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                lay(xd, eid, csr=csr)
            s3.record()
            for _ in range(20):
                lay(xd, eid, csr=csr)
            e3.record()
            torch.cuda.synchronize()
            return s3.elapsed_time(e3) / 20
    full_ms = layer_ms()                        # compact post path: [N, 4F] aggregate + pna_linear_scaled_fwd
    os.environ["PNA_B200_COMPACT_POST"] = "0"
    full_ms_12f = layer_ms()                    # the same layer through the materialised [N, 12F] tensor
    del os.environ["PNA_B200_COMPACT_POST"]

    cpu = None
    if not args.no_cpu_baseline:
        times, threads, info = cpu_reference_layer(ei, x, deg_hist, 3, 1)
        cpu = {"value": e * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": "port", **info,
               "sample": "full config-2 graph, 3 forward passes of PNAConvSimple's reference op sequence in torch CPU "
                         "(oracle/pna_oracle.py; torch_geometric/torch_scatter not installable)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "ogbn-arxiv-shaped CSR (BASELINE.json configs[1])", "n_nodes": n, "n_edges": e, "n_feat": f,
                   "aggregators": AGGRS, "scalers": SCALERS, "dst_skew": "perm[floor(N*u^3)]", "max_in_degree": csr.max_degree,
                   "split_rows": csr.n_hubs, "l2": "flushed between timed steps (512 MiB written, then 512 MiB read so no dirty lines remain)", "parallelism": "1 gpu"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "bytes_model": "B_min = N*F*s + 4E + 4(N+1) + 12*N*F*s",
                     "b_min_bytes": bytes_["b_min"], "b_gather_bytes": bytes_["b_gather"],
                     "effective_gbs_b_gather": bytes_["b_gather"] / (t_ms * 1e-3) / 1e9},
        "kernels_ms": {"step_min": min(per_step), "step_median": statistics.median(per_step),
                       "kernels": ("k_rows_stream (rows + chunks of split rows, split rows finalized by the last-arriving warp)"
                                   if fold_finalize_enabled() else "k_rows_stream (rows + chunks of split rows) + k_hub_finalize")
                                  + "; per-kernel times: profiles/"},
        "layer_fwd": {"ms": full_ms, "edges_per_s": e / (full_ms * 1e-3), "ms_via_12f_tensor": full_ms_12f,
                      "what": "PNAConvSimple.forward, CSR cached: aggregation with the identity scaler ([N,4F]) + post-MLP linear on "
                              "the tensor cores regenerating the scaled copies in registers (pna_linear_scaled_fwd, 3xTF32 "
                              "tcgen05); ms_via_12f_tensor = same layer through the materialised [N,12F] tensor"},
        "csr_build_ms": {"first_call": csr_ms_first, "steady": csr_ms},
        "e2e": {"value": e / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": x.numel() * 4 + ei.numel() * 8, "d2h_bytes_per_step": n * f * 4,
                "device_ms_per_step": s2.elapsed_time(e2) / k2, "wall_ms_per_step": e2e_wall_ms,
                "pcie_h2d_gbs": h2d_gbs, "pcie_d2h_gbs": d2h_gbs, "pinned": bool(xh.is_pinned() and outh.is_pinned()),
                "what": "PNAConvSimple.forward_host(x, edge_index) with pinned host tensors: H2D (x overlapped with the CSR build) + "
                        "aggregate + post-MLP in row blocks overlapped with the D2H of the result; CSR rebuilt every step"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
