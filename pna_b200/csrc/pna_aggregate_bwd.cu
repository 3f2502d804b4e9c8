// pna_aggregate_bwd: gradient of the aggregation w.r.t. the gathered rows (and the destination-side row_bias).
//
// Autograd of reference models/pytorch_geometric/aggregators.py:9-32 + scalers.py:8-29, as every training loop needs
// (multitask_benchmark/util/train.py:148, realworld_benchmark/train/*.py:31):
//   g_a   = sum over scalers s of scale_s(d) * grad_out[(s*A + a)*F + f]                        (per aggregator a)
//   d sum/dm = 1;  d mean/dm = 1/cnt;  d var/dm = 2 (m - mean)/cnt;  d std/dm = [var > 0] (m - mean)/(cnt * std)
//   min / max route to the FIRST slot attaining the extremum (torch_scatter's arg semantics)
// so that  grad_m(slot) = c0 + c1 * m + [slot == argmin] g_min + [slot == argmax] g_max  with two per-row coefficients.
// Two passes over the slots of a row: (A) recompute sum, sumsq, min, max and the first arg slots exactly as the forward
// does; (C) evaluate grad_m per slot, accumulate it into grad_gathered[col[slot]] (vector atomics -- several
// destinations share a source) and into grad_row_bias[row].  Rows at/above the split threshold are processed chunk by
// chunk by three small kernels (statistics, coefficients, scatter), like the forward.
//
// pna_aggregate_bwd_coef (gathered rows, col != NULL): pass C and its one vector atomic per (slot, feature chunk) are what
// bound the kernel above (REDG issue rate, ~0.9 TB/s of atomic traffic on the chip).  Per SOURCE row j the same gradient is
//   grad_gathered[j] = sum_{i <- j} (c0_i + c1_i * bias_i)  +  gathered[j] * sum_{i <- j} c1_i  +  routed min / max terms,
// so this entry point stops after the coefficients: it writes the row [c0' | c1] per destination, routes min / max with ONE
// scalar atomic per (row, feature) and writes grad_row_bias in closed form (deg * c0 + c1 * sum_m + gmin + gmax).  The two
// sums over the out-edges are a 'sum' aggregation of those rows over the transposed graph (pna_aggregate_fwd, no atomics),
// folded into grad_gathered by pna_aggregate_bwd_combine.
#include "pna_aggregate.cuh"
#include <string.h>

namespace pna {

struct BParams {
  KParams k;
  const void* go; long long ldgo;     // grad_out, layout of out
  float* gg; long long ldgg;          // grad_gathered [n_src, F] fp32, accumulated
  float* gb; long long ldgb;          // grad_row_bias [n_rows, F] fp32, written (nullable)
  int vec_atomics;                    // grad_gathered rows are 16-byte aligned
  float* coef; long long ldc;         // coefficient mode (pna_aggregate_bwd_coef): [n_rows, ldc] rows [c0' | c1], c1 at column coef_c1
  int coef_c1, coef_vec;
};

template <int VEC>
struct Stats {
  float sum[VEC], sq[VEC], mn[VEC], mx[VEC];
  int amn[VEC], amx[VEC];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sum[i] = 0.f; sq[i] = 0.f; mn[i] = CUDART_INF_F; mx[i] = -CUDART_INF_F; amn[i] = -1; amx[i] = -1; }
  }
  __device__ __forceinline__ void add(const float (&m)[VEC], int slot) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sum[i] = __fadd_rn(sum[i], m[i]);
      sq[i] = __fadd_rn(sq[i], __fmul_rn(m[i], m[i]));
      if (m[i] < mn[i]) { mn[i] = m[i]; amn[i] = slot; }   // strict: the first slot attaining the extremum wins
      if (m[i] > mx[i]) { mx[i] = m[i]; amx[i] = slot; }
    }
  }
};

template <int VEC>
struct Coef {
  float c0[VEC], c1[VEC], gmin[VEC], gmax[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ void load_m(const KParams& p, int src, int f, const float (&bias)[VEC], bool has_bias, float (&m)[VEC]) {
  Io<T, VEC>::load(local_row<T>(p, src) + f, m);
  if (has_bias) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = __fadd_rn(m[i], bias[i]);
  }
}

// upstream gradients of one row -> the two coefficients and the min / max gradients
template <typename T, int VEC>
__device__ __forceinline__ void coefficients(const BParams& b, long long row, int deg, int ooff, const Stats<VEC>& st, Coef<VEC>& c) {
  const KParams& p = b.k;
  const bool iso = deg == 0;
  const float degf = (float)deg, cnt = iso ? 1.0f : degf;
  const int sdeg = p.sdeg ? __ldg(p.sdeg + row) : deg;      // degree seen by the scalers (pna_agg_t.scaler_degree)
  const bool siso = sdeg == 0;
  const float sdegf = (float)sdeg;
  const float lg = logf(sdegf + 1.0f);
  const float s_amp = lg / p.avg_log, s_att = siso ? 1.0f : p.avg_log / lg;
  const float s_lin = sdegf / p.avg_lin, s_ilin = siso ? 1.0f : p.avg_lin / sdegf;
  const T* __restrict__ gorow = static_cast<const T*>(b.go) + row * b.ldgo + ooff;
#pragma unroll
  for (int i = 0; i < VEC; ++i) { c.c0[i] = 0.f; c.c1[i] = 0.f; c.gmin[i] = 0.f; c.gmax[i] = 0.f; }
  float mean[VEC], var[VEC], sd[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    mean[i] = st.sum[i] / cnt;
    var[i] = st.sq[i] / cnt - mean[i] * mean[i];
    sd[i] = sqrtf(fmaxf(var[i], 0.f) + 1e-5f);
  }
  for (int a = 0; a < p.nA; ++a) {
    const unsigned ac = (p.acodes >> (4 * a)) & 15u;
    if (ac == PNA_AGGR_SKIP) continue;
    float g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = 0.f;
    for (int s = 0; s < p.nS; ++s) {
      const unsigned sc = (p.scodes >> (4 * s)) & 15u;
      const float scale = sc == PNA_SCALE_IDENTITY ? 1.0f : sc == PNA_SCALE_AMPLIFICATION ? s_amp : sc == PNA_SCALE_ATTENUATION ? s_att
                          : sc == PNA_SCALE_LINEAR ? s_lin : s_ilin;
      float go[VEC];
      Io<T, VEC>::load(gorow + (s * p.nA + a) * p.Ft, go);
#pragma unroll
      for (int i = 0; i < VEC; ++i) g[i] += scale * go[i];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      switch (ac) {
        case PNA_AGGR_SUM: c.c0[i] += g[i]; break;
        case PNA_AGGR_MEAN: c.c0[i] += g[i] / cnt; break;
        case PNA_AGGR_MIN: c.gmin[i] += g[i]; break;
        case PNA_AGGR_MAX: c.gmax[i] += g[i]; break;
        case PNA_AGGR_VAR: {   // relu'(var) = [var > 0] in the DGL / dense flavours (PNA_FLAG_RELU_VAR)
          const float t = ((p.flags & PNA_FLAG_RELU_VAR) && !(var[i] > 0.f)) ? 0.f : 2.0f * g[i] / cnt;
          c.c1[i] += t; c.c0[i] -= t * mean[i];
        } break;
        default: { const float t = var[i] > 0.f ? g[i] / (cnt * sd[i]) : 0.f; c.c1[i] += t; c.c0[i] -= t * mean[i]; } break;
      }
    }
  }
}

template <int VEC>
__device__ __forceinline__ void scatter_grad(const BParams& b, long long src_row, int f, const float (&gm)[VEC], bool atomic) {
  float* dst = b.gg + src_row * b.ldgg + f;
  if (!atomic) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = gm[i];
    return;
  }
  if constexpr (VEC % 4 == 0) {
    if (b.vec_atomics) {
#pragma unroll
      for (int i = 0; i < VEC; i += 4) atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(gm[i], gm[i + 1], gm[i + 2], gm[i + 3]));
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) atomicAdd(dst + i, gm[i]);
}

// feature column / output column of this lane
struct LaneCols { int f, ooff; bool ok; };
template <int VEC>
__device__ __forceinline__ LaneCols lane_cols(const KParams& p, int gl, int fblock) {
  LaneCols c;
  c.f = fblock + gl * VEC;
  c.ok = c.f < p.F;
  if (!c.ok) { c.f = 0; }
  const int t = c.f / p.Ft, ft = c.f - t * p.Ft;
  c.ooff = t * p.Wt + p.has_self * p.Ft + ft;
  return c;
}

// coefficient mode: what one destination row hands to its sources.  c0' = c0 + c1 * bias; min / max go to the source of the
// one slot that attained them (st.amn / st.amx: absolute CSR slots, -1 = none); grad_row_bias in closed form.
template <int VEC>
__device__ __forceinline__ void emit_row(const BParams& b, long long row, int deg, const LaneCols& lc, const Stats<VEC>& st,
                                         const Coef<VEC>& c, const float (&bias)[VEC], bool has_bias) {
  const KParams& p = b.k;
  float* crow = b.coef + row * b.ldc + lc.f;
  float c0p[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) c0p[i] = has_bias ? fmaf(c.c1[i], bias[i], c.c0[i]) : c.c0[i];
  bool stored = false;
  if constexpr (VEC % 4 == 0) {
    if (b.coef_vec) {     // 16-byte aligned rows and halves
#pragma unroll
      for (int i = 0; i < VEC; i += 4) {
        *reinterpret_cast<float4*>(crow + i) = make_float4(c0p[i], c0p[i + 1], c0p[i + 2], c0p[i + 3]);
        *reinterpret_cast<float4*>(crow + b.coef_c1 + i) = make_float4(c.c1[i], c.c1[i + 1], c.c1[i + 2], c.c1[i + 3]);
      }
      stored = true;
    }
  }
  if (!stored) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { crow[i] = c0p[i]; crow[b.coef_c1 + i] = c.c1[i]; }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (c.gmin[i] != 0.f && st.amn[i] >= 0) atomicAdd(b.gg + (long long)__ldg(p.col + st.amn[i]) * b.ldgg + lc.f + i, c.gmin[i]);
    if (c.gmax[i] != 0.f && st.amx[i] >= 0) atomicAdd(b.gg + (long long)__ldg(p.col + st.amx[i]) * b.ldgg + lc.f + i, c.gmax[i]);
  }
  if (b.gb) {
    const float degf = (float)deg;
#pragma unroll
    for (int i = 0; i < VEC; ++i) b.gb[row * b.ldgb + lc.f + i] = degf * c.c0[i] + c.c1[i] * st.sum[i] + c.gmin[i] + c.gmax[i];
  }
}

constexpr int kBwdThreads = 256;

// ---- rows below the split threshold: one lane group per row --------------------------------------------------------
template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kBwdThreads) k_bwd_rows(const BParams b) {
  const KParams& p = b.k;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G;
  const long long row = ((long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (row >= p.n_rows) return;
  const LaneCols lc = lane_cols<VEC>(p, gl, blockIdx.y * (G * VEC));
  if (!lc.ok) return;
  const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1), deg = end - beg;
  if (deg >= p.split) return;
  const bool has_bias = p.bias != nullptr;
  float bias[VEC];
  if (has_bias) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + lc.f, bias);
  if (deg == 0) {
    if (b.gb) {
      float z[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) z[i] = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) b.gb[row * b.ldgb + lc.f + i] = z[i];
    }
    return;
  }
  constexpr int UB = 4;   // neighbour rows in flight per lane in each pass
  Stats<VEC> st;
  st.init();
  for (int e = beg; e < end; e += UB) {
    int src[UB];
    typename Io<T, VEC>::Raw raw[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) src[u] = (e + u < end) ? (p.col ? __ldg(p.col + e + u) : e + u) : -1;
#pragma unroll
    for (int u = 0; u < UB; ++u) if (src[u] >= 0) raw[u] = Io<T, VEC>::load_raw(local_row<T>(p, src[u]) + lc.f);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (src[u] >= 0) {
        float m[VEC];
        Io<T, VEC>::unpack(raw[u], m);
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) m[i] = __fadd_rn(m[i], bias[i]);
        }
        st.add(m, e + u);
      }
    }
  }
  Coef<VEC> c;
  coefficients<T, VEC>(b, row, deg, lc.ooff, st, c);
  if (b.coef) {   // coefficient mode: no second pass over the slots
    emit_row<VEC>(b, row, deg, lc, st, c, bias, has_bias);
    return;
  }
  float gbs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) gbs[i] = 0.f;
  for (int e = beg; e < end; e += UB) {
    int src[UB];
    typename Io<T, VEC>::Raw raw[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) src[u] = (e + u < end) ? (p.col ? __ldg(p.col + e + u) : e + u) : -1;
#pragma unroll
    for (int u = 0; u < UB; ++u) if (src[u] >= 0) raw[u] = Io<T, VEC>::load_raw(local_row<T>(p, src[u]) + lc.f);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (src[u] >= 0) {
        float m[VEC], gm[VEC];
        Io<T, VEC>::unpack(raw[u], m);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (has_bias) m[i] = __fadd_rn(m[i], bias[i]);
          gm[i] = c.c0[i] + c.c1[i] * m[i] + (e + u == st.amn[i] ? c.gmin[i] : 0.f) + (e + u == st.amx[i] ? c.gmax[i] : 0.f);
          gbs[i] += gm[i];
        }
        scatter_grad<VEC>(b, src[u], lc.f, gm, p.col != nullptr);
      }
    }
  }
  if (b.gb) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) b.gb[row * b.ldgb + lc.f + i] = gbs[i];
  }
}

// ---- split rows: chunk-parallel, like the forward ---------------------------------------------------------------
// (1) k_bwd_hub_stats: one lane group per 128-slot chunk -> partial sum, sumsq, min, max, first argmin / argmax slot;
// (2) k_bwd_hub_coef:  one lane group per split row merges its chunks in chunk order (strict < keeps the first slot)
//                      and turns the upstream gradient into (c0, c1, gmin, gmax, argmin, argmax) per feature;
// (3) k_bwd_hub_scatter: one lane group per chunk evaluates grad_m per slot, scatters it, and adds its share of the
//                      row_bias gradient.  Scratch (descriptor field hub_partials): 6*F floats per chunk + per split row.
template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kBwdThreads) k_bwd_hub_stats(const BParams b) {
  const KParams& p = b.k;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G;
  const long long c = ((long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (c >= p.n_chunks) return;
  const LaneCols lc = lane_cols<VEC>(p, gl, blockIdx.y * (G * VEC));
  if (!lc.ok) return;
  const int h = __ldg(p.chunk_items + 2 * c), j = __ldg(p.chunk_items + 2 * c + 1);
  const long long row = __ldg(p.hub_info + 4 * h);
  const int rbeg = __ldg(p.rowptr + row), rend = __ldg(p.rowptr + row + 1);
  const int beg = rbeg + j * p.chunk, end = min(beg + p.chunk, rend);
  const bool has_bias = p.bias != nullptr;
  float bias[VEC];
  if (has_bias) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + lc.f, bias);
  constexpr int UB = 4;
  Stats<VEC> st;
  st.init();
  for (int e = beg; e < end; e += UB) {
    int src[UB];
    typename Io<T, VEC>::Raw raw[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) src[u] = (e + u < end) ? (p.col ? __ldg(p.col + e + u) : e + u) : -1;
#pragma unroll
    for (int u = 0; u < UB; ++u) if (src[u] >= 0) raw[u] = Io<T, VEC>::load_raw(local_row<T>(p, src[u]) + lc.f);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (src[u] >= 0) {
        float m[VEC];
        Io<T, VEC>::unpack(raw[u], m);
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) m[i] = __fadd_rn(m[i], bias[i]);
        }
        st.add(m, e + u);
      }
    }
  }
  float* part = p.partials + c * 6ll * p.F + lc.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    part[0ll * p.F + i] = st.sum[i]; part[1ll * p.F + i] = st.sq[i]; part[2ll * p.F + i] = st.mn[i]; part[3ll * p.F + i] = st.mx[i];
    part[4ll * p.F + i] = __int_as_float(st.amn[i]); part[5ll * p.F + i] = __int_as_float(st.amx[i]);
  }
}

template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kBwdThreads) k_bwd_hub_coef(const BParams b) {
  const KParams& p = b.k;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G;
  const long long h = ((long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (h >= p.n_hubs) return;
  const LaneCols lc = lane_cols<VEC>(p, gl, blockIdx.y * (G * VEC));
  if (!lc.ok) return;
  const long long row = __ldg(p.hub_info + 4 * h);
  const int first = __ldg(p.hub_info + 4 * h + 1), nch = __ldg(p.hub_info + 4 * h + 2), deg = __ldg(p.hub_info + 4 * h + 3);
  Stats<VEC> st;
  st.init();
  for (int j = 0; j < nch; ++j) {
    const float* part = p.partials + (long long)(first + j) * 6ll * p.F + lc.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      st.sum[i] += part[0ll * p.F + i]; st.sq[i] += part[1ll * p.F + i];
      const float mn = part[2ll * p.F + i], mx = part[3ll * p.F + i];
      if (mn < st.mn[i]) { st.mn[i] = mn; st.amn[i] = __float_as_int(part[4ll * p.F + i]); }
      if (mx > st.mx[i]) { st.mx[i] = mx; st.amx[i] = __float_as_int(part[5ll * p.F + i]); }
    }
  }
  Coef<VEC> c;
  coefficients<T, VEC>(b, row, deg, lc.ooff, st, c);
  if (b.coef) {   // coefficient mode: the split row ends here, like every other row
    const bool has_bias = p.bias != nullptr;
    float bias[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) bias[i] = 0.f;
    if (has_bias) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + lc.f, bias);
    emit_row<VEC>(b, row, deg, lc, st, c, bias, has_bias);
    return;
  }
  float* co = p.partials + (p.n_chunks + h) * 6ll * p.F + lc.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    co[0ll * p.F + i] = c.c0[i]; co[1ll * p.F + i] = c.c1[i]; co[2ll * p.F + i] = c.gmin[i]; co[3ll * p.F + i] = c.gmax[i];
    co[4ll * p.F + i] = __int_as_float(st.amn[i]); co[5ll * p.F + i] = __int_as_float(st.amx[i]);
  }
  if (b.gb) {   // the chunks add their shares atomically in pass 3
#pragma unroll
    for (int i = 0; i < VEC; ++i) b.gb[row * b.ldgb + lc.f + i] = 0.f;
  }
}

template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kBwdThreads) k_bwd_hub_scatter(const BParams b) {
  const KParams& p = b.k;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G;
  const long long c = ((long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (c >= p.n_chunks) return;
  const LaneCols lc = lane_cols<VEC>(p, gl, blockIdx.y * (G * VEC));
  if (!lc.ok) return;
  const int h = __ldg(p.chunk_items + 2 * c), j = __ldg(p.chunk_items + 2 * c + 1);
  const long long row = __ldg(p.hub_info + 4 * h);
  const int rbeg = __ldg(p.rowptr + row), rend = __ldg(p.rowptr + row + 1);
  const int beg = rbeg + j * p.chunk, end = min(beg + p.chunk, rend);
  const bool has_bias = p.bias != nullptr;
  float bias[VEC];
  if (has_bias) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + lc.f, bias);
  const float* co = p.partials + (p.n_chunks + h) * 6ll * p.F + lc.f;
  Coef<VEC> cf;
  int amn[VEC], amx[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    cf.c0[i] = co[0ll * p.F + i]; cf.c1[i] = co[1ll * p.F + i]; cf.gmin[i] = co[2ll * p.F + i]; cf.gmax[i] = co[3ll * p.F + i];
    amn[i] = __float_as_int(co[4ll * p.F + i]); amx[i] = __float_as_int(co[5ll * p.F + i]);
  }
  constexpr int UB = 4;
  float gbs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) gbs[i] = 0.f;
  for (int e = beg; e < end; e += UB) {
    int src[UB];
    typename Io<T, VEC>::Raw raw[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) src[u] = (e + u < end) ? (p.col ? __ldg(p.col + e + u) : e + u) : -1;
#pragma unroll
    for (int u = 0; u < UB; ++u) if (src[u] >= 0) raw[u] = Io<T, VEC>::load_raw(local_row<T>(p, src[u]) + lc.f);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (src[u] >= 0) {
        float m[VEC], gm[VEC];
        Io<T, VEC>::unpack(raw[u], m);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (has_bias) m[i] = __fadd_rn(m[i], bias[i]);
          gm[i] = cf.c0[i] + cf.c1[i] * m[i] + (e + u == amn[i] ? cf.gmin[i] : 0.f) + (e + u == amx[i] ? cf.gmax[i] : 0.f);
          gbs[i] += gm[i];
        }
        scatter_grad<VEC>(b, src[u], lc.f, gm, p.col != nullptr);
      }
    }
  }
  if (b.gb) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomicAdd(b.gb + row * b.ldgb + lc.f + i, gbs[i]);
  }
}

template <typename T, int VEC, int G>
static int launch_bwd(const BParams& b, cudaStream_t st) {
  const KParams& p = b.k;
  constexpr int RPW = 32 / G;
  const unsigned gy = (unsigned)((p.F + G * VEC - 1) / (G * VEC));
  const long long per_block = (kBwdThreads / 32) * RPW;
  const long long gx = (p.n_rows + per_block - 1) / per_block;
  PNA_REQUIRE(gx <= 0x7fffffffll, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd: too many rows");
  k_bwd_rows<T, VEC, G><<<dim3((unsigned)gx, gy), kBwdThreads, 0, st>>>(b);
  PNA_CUDA_TRY(cudaGetLastError());
  if (p.n_hubs > 0) {
    const long long gc = (p.n_chunks + per_block - 1) / per_block, gh = (p.n_hubs + per_block - 1) / per_block;
    k_bwd_hub_stats<T, VEC, G><<<dim3((unsigned)gc, gy), kBwdThreads, 0, st>>>(b);
    PNA_CUDA_TRY(cudaGetLastError());
    k_bwd_hub_coef<T, VEC, G><<<dim3((unsigned)gh, gy), kBwdThreads, 0, st>>>(b);
    PNA_CUDA_TRY(cudaGetLastError());
    if (!b.coef) {
      k_bwd_hub_scatter<T, VEC, G><<<dim3((unsigned)gc, gy), kBwdThreads, 0, st>>>(b);
      PNA_CUDA_TRY(cudaGetLastError());
    }
  }
  return PNA_OK;
}

template <typename T, int VEC>
static int launch_bwd_typed(const BParams& b, cudaStream_t st) {
  const int chunks = b.k.F / VEC;
  if (chunks <= 1) return launch_bwd<T, VEC, 1>(b, st);
  if (chunks <= 2) return launch_bwd<T, VEC, 2>(b, st);
  if (chunks <= 4) return launch_bwd<T, VEC, 4>(b, st);
  if (chunks <= 8) return launch_bwd<T, VEC, 8>(b, st);
  if (chunks <= 16) return launch_bwd<T, VEC, 16>(b, st);
  return launch_bwd<T, VEC, 32>(b, st);     // wider rows: several feature blocks (gridDim.y)
}

static bool al16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

// ---- phase 3 of the coefficient path: grad_gathered[j] += S0[j] + gathered[j] * S1[j] ------------------------------------
// sums[j] = [S0 | S1] (S1 at column c1): the 'sum' of the coefficient rows over the out-edges of source row j.
template <typename T>
__global__ void __launch_bounds__(256) k_bwd_combine(const float* __restrict__ sums, long long lds, int c1, const T* __restrict__ x,
                                                     long long ldx, float* __restrict__ gg, long long ldgg, long long n_src, int F) {
  const long long total = n_src * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / F;
    const int f = (int)(i - r * F);
    const float s0 = sums[r * lds + f], s1 = sums[r * lds + c1 + f];
    float xv;
    if constexpr (sizeof(T) == 4) xv = x[r * ldx + f];
    else xv = __bfloat162float(x[r * ldx + f]);
    gg[r * ldgg + f] += s0 + xv * s1;
  }
}

}  // namespace pna

using namespace pna;

static int bwd_entry(const pna_agg_t* d, const void* grad_out, int64_t ld_grad_out, float* grad_gathered, int64_t ld_grad_gathered,
                     float* grad_row_bias, int64_t ld_grad_row_bias, float* coef, int64_t ld_coef, int32_t coef_c1,
                     pna_stream_t stream) {
  PNA_REQUIRE(d != nullptr, PNA_ERR_BAD_ARG, "pna_aggregate_bwd: null descriptor");
  PNA_REQUIRE(d->n_rows >= 0 && d->n_feat > 0 && d->n_towers > 0 && d->n_feat % d->n_towers == 0, PNA_ERR_BAD_ARG,
              "pna_aggregate_bwd: bad sizes");
  PNA_REQUIRE(d->n_aggr >= 1 && d->n_aggr <= PNA_MAX_AGGR && d->n_scalers >= 1 && d->n_scalers <= PNA_MAX_SCALERS, PNA_ERR_BAD_ARG,
              "pna_aggregate_bwd: n_aggr / n_scalers out of range");
  PNA_REQUIRE(d->dtype == PNA_F32 || d->dtype == PNA_BF16, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd: dtype %d", d->dtype);
  if (d->n_rows == 0) return PNA_OK;
  PNA_REQUIRE(d->gathered && d->rowptr && grad_out && grad_gathered, PNA_ERR_BAD_ARG, "pna_aggregate_bwd: null pointer");
  PNA_REQUIRE(d->peer_gathered == nullptr, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd: peer-memory graphs are forward-only");
  PNA_REQUIRE(d->ld_gathered < 0x3fffffffll, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd: row pitch too large");
  PNA_REQUIRE(d->split_threshold >= 2, PNA_ERR_BAD_ARG, "pna_aggregate_bwd: bad split threshold");
  if (d->n_hubs > 0)
    PNA_REQUIRE(d->hub_info && d->chunk_items && d->hub_partials && d->chunk_edges >= 1, PNA_ERR_BAD_ARG,
                "pna_aggregate_bwd: split rows need hub_info, chunk_items and hub_partials ((n_chunks + n_hubs) * 6 * n_feat floats)");

  BParams b;
  memset(&b, 0, sizeof(b));
  KParams& p = b.k;
  p.x = d->gathered; p.ldx = d->ld_gathered;
  p.rowptr = d->rowptr; p.col = d->col;
  p.bias = d->row_bias; p.ldb = d->ld_row_bias;
  p.n_rows = d->n_rows;
  p.F = d->n_feat; p.T = d->n_towers; p.Ft = d->n_feat / d->n_towers;
  p.has_self = d->self_feat ? 1 : 0;
  p.nA = d->n_aggr; p.nS = d->n_scalers; p.acodes = d->aggr_codes; p.scodes = d->scaler_codes;
  p.Wt = (p.has_self + p.nA * p.nS) * p.Ft;
  p.avg_log = d->avg_log; p.avg_lin = d->avg_lin;
  p.flags = d->flags; p.split = d->split_threshold; p.chunk = d->chunk_edges;
  p.hub_info = d->hub_info; p.n_hubs = d->n_hubs; p.chunk_items = d->chunk_items; p.n_chunks = d->n_chunks;
  p.partials = d->hub_partials;
  p.sdeg = d->scaler_degree;
  b.go = grad_out; b.ldgo = ld_grad_out;
  b.gg = grad_gathered; b.ldgg = ld_grad_gathered;
  b.gb = grad_row_bias; b.ldgb = ld_grad_row_bias;
  PNA_REQUIRE(b.ldgo >= (long long)p.T * p.Wt, PNA_ERR_BAD_ARG, "pna_aggregate_bwd: ld_grad_out too small");
  b.vec_atomics = al16(grad_gathered) && (ld_grad_gathered % 4 == 0);
  if (coef) {
    PNA_REQUIRE(d->col != nullptr, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd_coef: per-edge messages (col == NULL) have no shared sources");
    PNA_REQUIRE(coef_c1 >= p.F && ld_coef >= (long long)coef_c1 + p.F, PNA_ERR_BAD_ARG,
                "pna_aggregate_bwd_coef: coefficient rows need c1_column >= n_feat and ld_coef >= c1_column + n_feat");
    b.coef = coef; b.ldc = ld_coef; b.coef_c1 = coef_c1;
    b.coef_vec = al16(coef) && (ld_coef % 4 == 0) && (coef_c1 % 4 == 0);
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int esz = d->dtype == PNA_F32 ? 4 : 2;
  const int vec = 16 / esz;
  bool vec_ok = (p.Ft % vec == 0) && al16(p.x) && al16(grad_out) && (p.ldx % vec == 0) && (b.ldgo % vec == 0);
  if (p.bias) vec_ok = vec_ok && al16(p.bias) && (p.ldb % vec == 0);
  if (d->dtype == PNA_F32) return vec_ok ? launch_bwd_typed<float, 4>(b, st) : launch_bwd_typed<float, 1>(b, st);
  return vec_ok ? launch_bwd_typed<__nv_bfloat16, 8>(b, st) : launch_bwd_typed<__nv_bfloat16, 1>(b, st);
}

extern "C" int pna_aggregate_bwd(const pna_agg_t* d, const void* grad_out, int64_t ld_grad_out, float* grad_gathered,
                                 int64_t ld_grad_gathered, float* grad_row_bias, int64_t ld_grad_row_bias, pna_stream_t stream) {
  return bwd_entry(d, grad_out, ld_grad_out, grad_gathered, ld_grad_gathered, grad_row_bias, ld_grad_row_bias, nullptr, 0, 0, stream);
}

extern "C" int pna_aggregate_bwd_coef(const pna_agg_t* d, const void* grad_out, int64_t ld_grad_out, float* coef, int64_t ld_coef,
                                      int32_t c1_column, float* grad_gathered, int64_t ld_grad_gathered, float* grad_row_bias,
                                      int64_t ld_grad_row_bias, pna_stream_t stream) {
  PNA_REQUIRE(coef != nullptr, PNA_ERR_BAD_ARG, "pna_aggregate_bwd_coef: null coefficient buffer");
  return bwd_entry(d, grad_out, ld_grad_out, grad_gathered, ld_grad_gathered, grad_row_bias, ld_grad_row_bias, coef, ld_coef,
                   c1_column, stream);
}

extern "C" int pna_aggregate_bwd_combine(const float* coef_sums, int64_t ld_sums, int32_t c1_column, const void* gathered,
                                         int64_t ld_gathered, int32_t dtype, float* grad_gathered, int64_t ld_grad_gathered,
                                         int64_t n_src, int32_t n_feat, pna_stream_t stream) {
  PNA_REQUIRE(n_src >= 0 && n_feat > 0 && c1_column >= n_feat, PNA_ERR_BAD_ARG, "pna_aggregate_bwd_combine: bad sizes");
  PNA_REQUIRE(dtype == PNA_F32 || dtype == PNA_BF16, PNA_ERR_UNSUPPORTED, "pna_aggregate_bwd_combine: dtype %d", dtype);
  if (n_src == 0) return PNA_OK;
  PNA_REQUIRE(coef_sums && gathered && grad_gathered, PNA_ERR_BAD_ARG, "pna_aggregate_bwd_combine: null pointer");
  PNA_REQUIRE(ld_sums >= (long long)c1_column + n_feat && ld_gathered >= n_feat && ld_grad_gathered >= n_feat, PNA_ERR_BAD_ARG,
              "pna_aggregate_bwd_combine: row pitch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = n_src * (long long)n_feat;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;      // grid-stride: 16 CTAs of 256 threads per SM
  if (dtype == PNA_F32)
    k_bwd_combine<float><<<(unsigned)blocks, 256, 0, st>>>(coef_sums, ld_sums, c1_column, static_cast<const float*>(gathered),
                                                           ld_gathered, grad_gathered, ld_grad_gathered, n_src, n_feat);
  else
    k_bwd_combine<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(coef_sums, ld_sums, c1_column,
                                                                   static_cast<const __nv_bfloat16*>(gathered), ld_gathered,
                                                                   grad_gathered, ld_grad_gathered, n_src, n_feat);
  PNA_CUDA_TRY(cudaGetLastError());
  return PNA_OK;
}
