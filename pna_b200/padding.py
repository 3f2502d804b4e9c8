"""Feature widths that are not a multiple of 16 bytes (ZINC: 75 = 5 towers x 15; superpixel nets: 70) would force the
kernel onto its 4- / 2-byte scalar path.  Inside the layers the padded width costs nothing instead:

* the gathered rows are produced directly at the padded width (zero columns appended to x, or zero rows appended to the
  pre-transformation weights, so the GEMM writes the pad columns as exact zeros);
* the aggregation runs on the 128-bit path and its output keeps the padded column blocks;
* the first post-transformation Linear gets zero weight columns at the pad positions, so the pad features (0, or the
  constant sqrt(1e-5) of the std columns) contribute exactly 0 to the result.
No large tensor is copied or re-laid out; only [out, k*F] weight matrices are expanded per call.
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


def vec_elems(dtype: torch.dtype) -> int:
    return 16 // torch.empty(0, dtype=dtype).element_size()


def padded_width(f: int, dtype: torch.dtype) -> int:
    v = vec_elems(dtype)
    return (f + v - 1) // v * v


def pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    """[N, f] -> [N, width] with zero columns appended."""
    return t if t.size(-1) == width else Fn.pad(t, (0, width - t.size(-1)))


def pad_blocks(t: torch.Tensor, n_blocks: int, f: int, fp: int) -> torch.Tensor:
    """[N, n_blocks*f] -> [N, n_blocks*fp]: every block of f columns gets fp - f zero columns appended."""
    if f == fp:
        return t
    return Fn.pad(t.reshape(t.size(0), n_blocks, f), (0, fp - f)).reshape(t.size(0), n_blocks * fp)


def expand_weight_cols(w: torch.Tensor, n_blocks: int, f: int, fp: int) -> torch.Tensor:
    """Linear weight [out, n_blocks*f] -> [out, n_blocks*fp] with zero columns at the pad positions."""
    if f == fp:
        return w
    return Fn.pad(w.reshape(w.size(0), n_blocks, f), (0, fp - f)).reshape(w.size(0), n_blocks * fp)


def expand_weight_rows(w: torch.Tensor, f: int, fp: int) -> torch.Tensor:
    """Linear weight [f, in] -> [fp, in] with zero rows appended (the GEMM then emits zero pad features)."""
    return w if f == fp else Fn.pad(w, (0, 0, 0, fp - f))
