"""Layer-output error of PNAConvSimple.forward on the GPU and of the CPU reference op sequence, both against float64."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pna_b200
from pna_b200 import synth
from oracle import pna_oracle as O
A4, S3 = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
dev = torch.device("cuda:0")
res = []
for name, (ei, x) in (("config2 F=128 (K=1536)", synth.arxiv_like()), ("smoke F=128 n=3000", synth.arxiv_like(n_nodes=3000, n_edges=30000, seed=1))):
    n, f = x.shape
    deg = synth.degree_histogram(ei[1], n)
    ref = O.PNAConvSimpleOracle(f, f, A4, S3, deg)
    lay = pna_b200.PNAConvSimple(f, f, A4, S3, deg); lay.load_state_dict(ref.state_dict()); lay = lay.to(dev)
    with torch.no_grad():
        got = lay(x.to(dev), ei.to(dev)).cpu()
        os.environ["PNA_B200_TENSOR_LINEAR"] = "0"
        got_cublas = lay(x.to(dev), ei.to(dev)).cpu()
        del os.environ["PNA_B200_TENSOR_LINEAR"]
        want32 = ref(x, ei)
        agg64 = O.simple_propagate(x.double(), ei, A4, S3, ref.avg_deg)
        W, b = ref.post_nn[0].weight.double(), ref.post_nn[0].bias.double()
        want64 = agg64 @ W.t() + b
        cond = agg64.abs() @ W.abs().t() + b.abs()
    e = lambda a: dict(max=float((a.double() - want64).abs().max()), rms=float((a.double() - want64).pow(2).mean().sqrt()),
                       max_over_cond=float(((a.double() - want64).abs() / cond).max()))
    res.append({"case": name, "gpu_3xtf32": e(got), "gpu_cublas_fp32": e(got_cublas), "cpu_reference_fp32": e(want32),
                "gpu_vs_cpu32_max": float((got - want32).abs().max()), "out_abs_max": float(want64.abs().max()),
                "out_abs_mean": float(want64.abs().mean())})
print(json.dumps(res, indent=1))
