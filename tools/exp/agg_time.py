"""Time pna_aggregate_fwd on one BASELINE config shape; parity of sampled rows vs the oracle.  Tuning experiments.

    python tools/exp/agg_time.py --config 2 [--steps 30] [--check 4000] [--once]
--once: one warm-up + two calls only (for runs under ncu)."""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pna_b200
from pna_b200 import synth
from oracle import pna_oracle as O

A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="2")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--check", type=int, default=4000)
ap.add_argument("--once", action="store_true")
ap.add_argument("--tag", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
if os.environ.get("PNA_EXP_L2_PERSIST_MB"):     # experiment: L2 set-aside for persisting (evict_last) lines
    import ctypes
    torch.cuda.init(); torch.zeros(1, device=dev)
    rt = ctypes.CDLL("libcudart.so.12")
    mb = int(os.environ["PNA_EXP_L2_PERSIST_MB"])
    rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(mb << 20))      # cudaLimitPersistingL2CacheSize
    got = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(got), 6)
    print(f"# persisting L2 limit: asked {mb} MiB rc {rc} -> {got.value >> 20} MiB", file=sys.stderr)
peak = 6571.6
pk = os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def make(cfg):
    if cfg == "1": return synth.multitask_like()
    if cfg == "2": return synth.arxiv_like()
    if cfg == "2u": return synth.arxiv_like(skew=1.0)
    if cfg == "3": return synth.zinc_like(dtype=torch.bfloat16)[:2]
    if cfg == "3p": return synth.zinc_like(n_feat=80, dtype=torch.bfloat16)[:2]
    if cfg == "3f": return synth.zinc_like(dtype=torch.float32)[:2]
    if cfg == "4": return synth.superpixel_like()
    if cfg == "5": return synth.powerlaw()
    if cfg == "5u":   # power-law destinations, UNIFORM sources: no hot source rows
        ei, x = synth.powerlaw()
        g = torch.Generator().manual_seed(5)
        return torch.stack([torch.randint(0, x.size(0), (ei.size(1),), generator=g), ei[1]]), x
    if cfg == "5w":   # the same rows with (almost) no edges: the write path alone
        ei, x = synth.powerlaw()
        return ei[:, :1000], x
    raise SystemExit("unknown config")


ei, x = make(args.config)
n, f = x.shape
e = ei.size(1)
avg = pna_b200.avg_deg_from_histogram(synth.degree_histogram(ei[1], n))
xd = x.to(dev)
csr = pna_b200.build_csr(ei[0].to(dev), ei[1].to(dev), n)
out = torch.empty((n, 12 * f), dtype=x.dtype, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(128 << 20, dtype=torch.float32, device=dev)


def step():
    pna_b200.aggregate_forward(xd, csr, A4, S3, avg, out=out)


if args.once:
    step(); torch.cuda.synchronize()
    flush.zero_(); flush_rd.sum(); step(); torch.cuda.synchronize()
    flush.zero_(); flush_rd.sum(); step(); torch.cuda.synchronize()
    sys.exit(0)

ts = []
for i in range(args.steps + 5):
    flush.zero_(); flush_rd.sum()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); step(); t.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(s.elapsed_time(t))
ms = statistics.mean(ts)
# parity on randomly sampled rows (all of their in-edges), light rows at 1e-5 / bf16 tolerance, split rows vs float64
g = torch.Generator().manual_seed(7)
rows = torch.randperm(n, generator=g)[: min(args.check, n)]
mark = torch.zeros(n, dtype=torch.bool); mark[rows] = True
keep = mark[ei[1]]
sub = ei[:, keep]
want = O.simple_propagate(x.float(), sub, A4, S3, avg)[rows]
got = out[rows.to(dev)].float().cpu()
deg = torch.bincount(sub[1], minlength=n)[rows]
light = deg < csr.split_threshold
tol = dict(rtol=1e-5, atol=1e-5) if x.dtype == torch.float32 else dict(rtol=2 ** -8, atol=1e-3)
ok = torch.allclose(got[light], want[light], **tol)
err = (got[light] - want[light]).abs().max().item() if light.any() else 0.0
hub_err = ((got[~light] - want[~light]).abs() / (1 + want[~light].abs())).max().item() if (~light).any() else 0.0
by = synth.algorithmic_bytes(n, e, f, x.element_size(), 12 * f)
print(json.dumps({"tag": args.tag, "config": args.config, "env": {k: v for k, v in os.environ.items() if k.startswith("PNA_B200")},
                  "n": n, "e": e, "f": f, "ms_mean": ms, "ms_min": min(ts), "ms_median": statistics.median(ts),
                  "frac": by["b_min"] / ms / 1e6 / peak, "parity_light": bool(ok), "max_err_light": err, "hub_rel_err": hub_err,
                  "hubs": csr.n_hubs, "max_deg": csr.max_degree}), flush=True)
