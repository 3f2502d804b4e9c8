"""fp32 error of the two backward paths against the oracle's float64 autograd on a power-law multigraph, both executed on the
HOST through tests/emu (no GPU needed).  Numbers in profiles/r02_backward_ab.json.

    python tools/exp/bwd_accuracy.py 60000 2000000"""
import ctypes as C, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from pna_b200 import _lib, synth
import test_bwd_emulated as T
from oracle import pna_oracle as O

spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "emu", "build_emu.py"))
be = importlib.util.module_from_spec(spec); spec.loader.exec_module(be)
L = C.CDLL(be.build()); L.emu_last_error.restype = C.c_char_p
L.pna_aggregate_bwd.argtypes = [C.POINTER(_lib.AggStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
L.pna_aggregate_bwd_coef.argtypes = [C.POINTER(_lib.AggStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int64, C.c_void_p]
L.pna_aggregate_bwd_combine.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int32, C.c_void_p]
T.SPLIT, T.CHUNK = 256, 128                        # the library's defaults
n, e, f = int(sys.argv[1]), int(sys.argv[2]), 8
ei, _ = synth.powerlaw(n_nodes=n, n_edges=e, n_feat=f, with_features=False)
src, dst = ei[0], ei[1]
x = synth.hash_features(torch.arange(n), f)
aggrs, scalers = ["mean", "std"], ["identity", "amplification", "attenuation"]      # no min / max: torch's CPU backward splits ties
avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(dst, minlength=n)))
w = torch.randn(n, len(aggrs) * len(scalers) * f, generator=torch.Generator().manual_seed(1))
(g_atomic, _), (g_coef, _) = T.run_both(L, x, None, src, dst, n, w, aggrs, scalers, avg)
xr = x.double().clone().requires_grad_(True)
(O.pyg_aggregate(xr[src], dst, n, aggrs, scalers, avg) * w.double()).sum().backward()
truth = xr.grad
x32 = x.clone().requires_grad_(True)
(O.pyg_aggregate(x32[src], dst, n, aggrs, scalers, avg) * w).sum().backward()
for name, g in (("reference fp32 autograd", x32.grad), ("atomic (per-edge evaluation)", g_atomic), ("coef (regrouped per source)", g_coef)):
    d = (g.double() - truth).abs()
    r = int(d.argmax()) // f
    print(f"{name:32s} rel Frobenius error {float(d.norm() / truth.norm()):.2e}  max abs {float(d.max()):.3g} at a row with "
          f"{int((src == r).sum())} out-edges (max |grad| {float(truth.abs().max()):.3g})")
