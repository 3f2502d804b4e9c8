"""ctypes loader of the plain-C oracle (oracle/c/pna_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpna_oracle.so")
AGGR = {"sum": 0, "mean": 1, "min": 2, "max": 3, "var": 4, "std": 5}
SCAL = {"identity": 0, "amplification": 1, "attenuation": 2, "linear": 3, "inverse_linear": 4}


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


def aggregate(x, edge_index, aggregators, scalers, avg_deg, zero_isolated=False):
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    x = x.contiguous().float()
    n, f = x.shape
    src, dst = edge_index[0].contiguous().long(), edge_index[1].contiguous().long()
    a = (C.c_int32 * len(aggregators))(*[AGGR[k] for k in aggregators])
    s = (C.c_int32 * len(scalers))(*[SCAL[k] for k in scalers])
    out = torch.empty((n, len(aggregators) * len(scalers) * f), dtype=torch.float32)
    rc = lib.pna_oracle_aggregate(
        C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(f), C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
        C.c_int64(src.numel()), a, len(aggregators), s, len(scalers), C.c_float(avg_deg["log"]),
        C.c_float(avg_deg.get("lin", 1.0)), int(zero_isolated), C.c_void_p(out.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"pna_oracle_aggregate failed: {rc}")
    return out
