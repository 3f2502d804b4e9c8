// pna_aggregate_fwd: kernels + host dispatch.  See pna_aggregate.cuh for the design notes.
#include "pna_aggregate.cuh"

namespace pna {

constexpr int kThreads = 256;  // 8 warps per CTA
// Resident CTAs per SM the register allocator must leave room for (latency hiding for the gather needs warps):
// 16 accumulators per 128-bit chunk per lane bound what is possible.
constexpr int min_blocks(int vec, int k) { return vec * k <= 4 ? 4 : (vec * k <= 8 ? 3 : (vec * k <= 16 ? 2 : 1)); }

// fp32 partials of a chunk: default cache policy (re-read by k_hub_finalize right after).
template <int VEC>
__device__ __forceinline__ void store_f32(float* p, const float (&v)[VEC]) {
  if constexpr (VEC % 4 == 0) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = v[i];
  }
}

// ---- rows below the split threshold: one lane group per row ------------------------------------------------
template <typename T, int VEC, int G, int K, int U>
__global__ void __launch_bounds__(kThreads, min_blocks(VEC, K)) k_rows(const KParams p) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const long long slot = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  long long row;
  if (p.row_ids) {
    if (slot >= p.n_row_ids) return;
    row = __ldg(p.row_ids + slot);
  } else {
    if (slot >= p.n_rows) return;
    row = slot;
  }
  const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
  const int deg = end - beg;
  if (deg >= p.split) return;  // hubs are handled by k_hub_chunks + k_hub_finalize

  FeatMap<VEC, G, K> fm;
  fm.init(p, gl, blockIdx.y * (G * VEC * K));
  Acc<VEC> acc[K];
  float bias[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    acc[k].init();
    if (p.bias && fm.ok[k]) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + fm.f[k], bias[k]);
  }
  accumulate_slots<T, VEC, G, K, U>(p, fm, beg, end, bias, p.bias != nullptr, acc);
  finalize_row<T, VEC, G, K>(p, fm, row, deg, acc);
}

// ---- hubs, pass 1: one lane group per chunk of `chunk` slots -> fp32 partials ------------------------------
template <typename T, int VEC, int G, int K, int U>
__global__ void __launch_bounds__(kThreads, min_blocks(VEC, K)) k_hub_chunks(const KParams p) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const long long c = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (c >= p.n_chunks) return;
  const int h = __ldg(p.chunk_items + 2 * c), j = __ldg(p.chunk_items + 2 * c + 1);
  const long long row = __ldg(p.hub_info + 4 * h);
  const int rbeg = __ldg(p.rowptr + row), rend = __ldg(p.rowptr + row + 1);
  const int beg = rbeg + j * p.chunk;
  const int end = min(beg + p.chunk, rend);

  FeatMap<VEC, G, K> fm;
  fm.init(p, gl, blockIdx.y * (G * VEC * K));
  Acc<VEC> acc[K];
  float bias[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    acc[k].init();
    if (p.bias && fm.ok[k]) Io<T, VEC>::load(static_cast<const T*>(p.bias) + row * p.ldb + fm.f[k], bias[k]);
  }
  accumulate_slots<T, VEC, G, K, U>(p, fm, beg, end, bias, p.bias != nullptr, acc);

  float* __restrict__ part = p.partials + c * 4ll * p.F;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (!fm.ok[k]) continue;
    store_f32<VEC>(part + 0ll * p.F + fm.f[k], acc[k].sum);
    store_f32<VEC>(part + 1ll * p.F + fm.f[k], acc[k].sq);
    store_f32<VEC>(part + 2ll * p.F + fm.f[k], acc[k].mn);
    store_f32<VEC>(part + 3ll * p.F + fm.f[k], acc[k].mx);
  }
}

// ---- hubs, pass 2: merge a hub's partials in chunk order, then the common epilogue -------------------------
template <typename T, int VEC, int G, int K>
__global__ void __launch_bounds__(kThreads) k_hub_finalize(const KParams p) {
  constexpr int RPW = 32 / G;
  constexpr int UF = (K * VEC >= 16) ? 1 : (K * VEC >= 8 ? 2 : 4);  // partial sets in flight (register budget)
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const long long h = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (h >= p.n_hubs) return;
  const long long row = __ldg(p.hub_info + 4 * h);
  const int first = __ldg(p.hub_info + 4 * h + 1), nch = __ldg(p.hub_info + 4 * h + 2);
  const int deg = __ldg(p.hub_info + 4 * h + 3);

  FeatMap<VEC, G, K> fm;
  fm.init(p, gl, blockIdx.y * (G * VEC * K));
  Acc<VEC> acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k].init();

  for (int j = 0; j < nch; j += UF) {
    float ps[UF][K][4][VEC];
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      if (j + u < nch) {
        const float* part = p.partials + (long long)(first + j + u) * 4ll * p.F;
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (fm.ok[k]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // plain (coherent) loads: the partials were written by the previous kernel on this stream
              const float* src = part + (long long)q * p.F + fm.f[k];
#pragma unroll
              for (int i = 0; i < VEC; ++i) ps[u][k][q][i] = src[i];
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      if (j + u < nch) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (fm.ok[k]) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              acc[k].sum[i] = __fadd_rn(acc[k].sum[i], ps[u][k][0][i]);
              acc[k].sq[i] = __fadd_rn(acc[k].sq[i], ps[u][k][1][i]);
              acc[k].mn[i] = fminf(acc[k].mn[i], ps[u][k][2][i]);
              acc[k].mx[i] = fmaxf(acc[k].mx[i], ps[u][k][3][i]);
            }
          }
      }
    }
  }
  finalize_row<T, VEC, G, K>(p, fm, row, deg, acc);
}

// ---- host dispatch -----------------------------------------------------------------------------------------
template <typename T, int VEC, int G, int K, int U>
static int launch_config(const KParams& p, cudaStream_t st) {
  constexpr int RPW = 32 / G;
  constexpr int per_block = (kThreads / 32) * RPW;
  const unsigned gy = (unsigned)((p.F + G * VEC * K - 1) / (G * VEC * K));
  if (!(p.flags & PNA_FLAG_SKIP_LIGHT)) {
    const long long slots = p.row_ids ? p.n_row_ids : p.n_rows;
    if (slots > 0) {
      const long long gx = (slots + per_block - 1) / per_block;
      PNA_REQUIRE(gx <= 0x7fffffffll, PNA_ERR_UNSUPPORTED, "too many rows for one launch: %lld", slots);
      k_rows<T, VEC, G, K, U><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>(p);
      PNA_CUDA_TRY(cudaGetLastError());
    }
  }
  if (!(p.flags & PNA_FLAG_SKIP_HUBS) && p.n_hubs > 0) {
    const long long gc = (p.n_chunks + per_block - 1) / per_block;
    k_hub_chunks<T, VEC, G, K, U><<<dim3((unsigned)gc, gy), kThreads, 0, st>>>(p);
    PNA_CUDA_TRY(cudaGetLastError());
    const long long gh = (p.n_hubs + per_block - 1) / per_block;
    k_hub_finalize<T, VEC, G, K><<<dim3((unsigned)gh, gy), kThreads, 0, st>>>(p);
    PNA_CUDA_TRY(cudaGetLastError());
  }
  return PNA_OK;
}

template <typename T, int VEC>
static int launch_typed(const KParams& p, cudaStream_t st) {
  const int chunks = p.F / VEC;  // VEC divides F on the vector path; VEC == 1 on the scalar path
  if (chunks <= 1) return launch_config<T, VEC, 1, 1, 4>(p, st);
  if (chunks <= 2) return launch_config<T, VEC, 2, 1, 4>(p, st);
  if (chunks <= 4) return launch_config<T, VEC, 4, 1, 4>(p, st);
  if (chunks <= 8) return launch_config<T, VEC, 8, 1, 4>(p, st);
  if (chunks <= 16) return launch_config<T, VEC, 16, 1, 4>(p, st);
  if (chunks <= 32) return launch_config<T, VEC, 32, 1, 4>(p, st);
  if (chunks <= 64) return launch_config<T, VEC, 32, 2, 2>(p, st);
  if (chunks <= 96) return launch_config<T, VEC, 32, 3, 2>(p, st);
  return launch_config<T, VEC, 32, 4, 2>(p, st);  // wider rows: several feature blocks (gridDim.y)
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace pna

using namespace pna;

extern "C" int pna_aggregate_fwd(const pna_agg_t* d, pna_stream_t stream) {
  PNA_REQUIRE(d != nullptr, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: null descriptor");
  PNA_REQUIRE(d->n_rows >= 0 && d->n_feat > 0 && d->n_towers > 0, PNA_ERR_BAD_ARG,
              "pna_aggregate_fwd: bad sizes n_rows=%lld n_feat=%d n_towers=%d", (long long)d->n_rows, d->n_feat,
              d->n_towers);
  PNA_REQUIRE(d->n_feat % d->n_towers == 0, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: n_feat %d not divisible by n_towers %d",
              d->n_feat, d->n_towers);
  PNA_REQUIRE(d->n_feat <= pna_query(PNA_QUERY_MAX_FEATURES), PNA_ERR_UNSUPPORTED, "pna_aggregate_fwd: n_feat %d too large",
              d->n_feat);
  PNA_REQUIRE(d->n_aggr >= 1 && d->n_aggr <= PNA_MAX_AGGR && d->n_scalers >= 1 && d->n_scalers <= PNA_MAX_SCALERS,
              PNA_ERR_BAD_ARG, "pna_aggregate_fwd: n_aggr=%d n_scalers=%d out of range", d->n_aggr, d->n_scalers);
  for (int a = 0; a < d->n_aggr; ++a)
    PNA_REQUIRE(((d->aggr_codes >> (4 * a)) & 15u) <= PNA_AGGR_STD, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: bad aggregator code");
  for (int s = 0; s < d->n_scalers; ++s)
    PNA_REQUIRE(((d->scaler_codes >> (4 * s)) & 15u) <= PNA_SCALE_INVERSE_LINEAR, PNA_ERR_BAD_ARG,
                "pna_aggregate_fwd: bad scaler code");
  PNA_REQUIRE(d->dtype == PNA_F32 || d->dtype == PNA_BF16, PNA_ERR_UNSUPPORTED, "pna_aggregate_fwd: dtype %d", d->dtype);
  if (d->n_rows == 0) return PNA_OK;
  PNA_REQUIRE(d->gathered && d->rowptr && d->out, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: null gathered/rowptr/out");
  PNA_REQUIRE(d->split_threshold >= 2 && d->chunk_edges >= 1, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: bad split/chunk");
  if (d->n_hubs > 0 && !(d->flags & PNA_FLAG_SKIP_HUBS))
    PNA_REQUIRE(d->hub_info && d->chunk_items && d->hub_partials, PNA_ERR_BAD_ARG,
                "pna_aggregate_fwd: n_hubs=%lld but hub_info/chunk_items/hub_partials missing", (long long)d->n_hubs);

  KParams p;
  p.x = d->gathered; p.ldx = d->ld_gathered;
  p.rowptr = d->rowptr; p.col = d->col;
  p.bias = d->row_bias; p.ldb = d->ld_row_bias;
  p.self = d->self_feat; p.lds = d->ld_self; p.self_tstride = d->self_tower_stride;
  p.out = d->out; p.ldo = d->ld_out;
  p.n_rows = d->n_rows;
  p.F = d->n_feat; p.T = d->n_towers; p.Ft = d->n_feat / d->n_towers;
  p.has_self = d->self_feat ? 1 : 0;
  p.nA = d->n_aggr; p.nS = d->n_scalers; p.acodes = d->aggr_codes; p.scodes = d->scaler_codes;
  p.Wt = (p.has_self + p.nA * p.nS) * p.Ft;
  p.avg_log = d->avg_log; p.avg_lin = d->avg_lin;
  p.flags = d->flags; p.split = d->split_threshold; p.chunk = d->chunk_edges;
  p.hub_info = d->hub_info; p.chunk_items = d->chunk_items; p.n_hubs = d->n_hubs; p.n_chunks = d->n_chunks;
  p.partials = d->hub_partials;
  p.row_ids = d->row_ids; p.n_row_ids = d->row_ids ? d->n_row_ids : 0;
  PNA_REQUIRE(p.ldo >= (long long)p.T * p.Wt, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: ld_out %lld < row width %lld",
              (long long)p.ldo, (long long)p.T * p.Wt);

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int esz = d->dtype == PNA_F32 ? 4 : 2;
  const int vec = 16 / esz;
  // 128-bit path needs every row start and every output segment 16-byte aligned.
  bool vec_ok = (p.Ft % vec == 0) && aligned16(p.x) && aligned16(p.out) && (p.ldx % vec == 0) && (p.ldo % vec == 0);
  if (p.bias) vec_ok = vec_ok && aligned16(p.bias) && (p.ldb % vec == 0);
  if (p.self) vec_ok = vec_ok && aligned16(p.self) && (p.lds % vec == 0) && (p.self_tstride % vec == 0);
  if (d->dtype == PNA_F32) {
    return vec_ok ? launch_typed<float, 4>(p, st) : launch_typed<float, 1>(p, st);
  } else {
    return vec_ok ? launch_typed<__nv_bfloat16, 8>(p, st) : launch_typed<__nv_bfloat16, 1>(p, st);
  }
}
