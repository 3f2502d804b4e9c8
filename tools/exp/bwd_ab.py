"""Backward of the aggregation, the two paths side by side (PNA_B200_BWD=atomic | coef), at config 2 and the config-5 share.
Prints one JSON line per config: ms of each path (CUDA events, median), the forward beside them, and how far the two
gradients are apart.

    python tools/exp/bwd_ab.py [--configs 2,5] [--steps 5]"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pna_b200
from pna_b200 import synth
from pna_b200.aggregate import aggregate_backward, aggregate_forward

A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="2,5")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda:0")


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for cfg in args.configs.split(","):
    if cfg == "2":
        ei, x = synth.arxiv_like()
        src, dst, x = ei[0].to(dev), ei[1].to(dev), x.to(dev)
    else:
        n, e, f = 1_250_000, 12_500_000, 256
        src, dst = next(synth.powerlaw_stream(n, e, dev))
        x = synth.hash_features(torch.arange(n, device=dev), f)
    n, f = x.shape
    csr = pna_b200.build_csr(src, dst, n)
    hist = torch.bincount(csr.in_degree.long()).cpu()
    avg = pna_b200.avg_deg_from_histogram(hist)
    go = torch.randn((n, 12 * f), device=dev)
    out = torch.empty((n, 12 * f), device=dev)
    res = {"config": cfg, "n_nodes": n, "n_edges": int(src.numel()), "n_feat": f, "split_rows": csr.n_hubs, "max_in_degree": csr.max_degree}
    res["forward_ms"] = timed(lambda: aggregate_forward(x, csr, A4, S3, avg, out=out), args.steps)
    grads = {}
    for mode in ("atomic", "coef"):
        os.environ["PNA_B200_BWD"] = mode
        steps = args.steps if not (mode == "atomic" and cfg != "2") else max(2, args.steps // 2)
        res[f"backward_{mode}_ms"] = timed(lambda: aggregate_backward(go, x, csr, A4, S3, avg), steps, warm=1 if cfg != "2" else 2)
        grads[mode] = aggregate_backward(go, x, csr, A4, S3, avg)[0]
    torch.cuda.synchronize()
    d = (grads["coef"] - grads["atomic"]).abs()
    res["max_abs_diff"] = float(d.max())
    res["max_abs_grad"] = float(grads["atomic"].abs().max())
    res["rel_fro_diff"] = float(d.norm() / grads["atomic"].norm())
    res["backward_over_forward"] = {m: res[f"backward_{m}_ms"] / res["forward_ms"] for m in ("atomic", "coef")}
    print(json.dumps(res), flush=True)
    del go, out, grads, csr, x
    torch.cuda.empty_cache()
