// pna_aggregate_fwd kernels + launch dispatch (templates).  Instantiated per dtype / vector width in
// pna_aggregate_{f32,bf16}_{vec,scalar}.cu so the translation units compile in parallel.
#pragma once
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
template <typename T, int VEC, int G, int K, int U, typename Cfg>
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
  finalize_row<T, VEC, G, K, Cfg>(p, fm, row, deg, acc);
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

// ---- hubs, pass 2: one CTA per hub merges its partials, then the common epilogue ---------------------------
// kFinGroups lane groups stride over the hub's chunks (group q takes chunks q, q+kFinGroups, ..), each merging in
// chunk order; the per-group results are parked in the groups' own first partial slots (scratch, rebuilt every call)
// and group 0 merges those in group order.  Deterministic; no atomics; a 20k-edge hub is ~20 chunk reads per group.
constexpr int kFinGroups = 8;

template <int VEC, int K, int UF>
__device__ __forceinline__ void merge_partials(const float* __restrict__ base, long long F, const int (&f)[K], const bool (&ok)[K],
                                               int first, int count, int stride, Acc<VEC> (&acc)[K]) {
  for (int j = 0; j < count; j += UF) {
    float ps[UF][K][4][VEC];
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      if (j + u < count) {
        const float* part = base + (long long)(first + (long long)(j + u) * stride) * 4ll * F;
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (ok[k]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float* src = part + (long long)q * F + f[k];   // plain loads: written earlier on this stream / CTA
              if constexpr (VEC % 4 == 0) {
#pragma unroll
                for (int i = 0; i < VEC; i += 4) {
                  const float4 t = *reinterpret_cast<const float4*>(src + i);
                  ps[u][k][q][i] = t.x; ps[u][k][q][i + 1] = t.y; ps[u][k][q][i + 2] = t.z; ps[u][k][q][i + 3] = t.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) ps[u][k][q][i] = src[i];
              }
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      if (j + u < count) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (ok[k]) {
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
}

template <typename T, int VEC, int G, int K>
__global__ void __launch_bounds__(kFinGroups * 32) k_hub_finalize(const KParams p) {
  constexpr int UF = (K * VEC >= 16) ? 1 : (K * VEC >= 8 ? 2 : 4);  // partial sets in flight (register budget)
  const int gl = threadIdx.x % G;
  const int q = threadIdx.x / G;          // lane group within the CTA, 0..kFinGroups-1
  const long long h = blockIdx.x;
  const long long row = __ldg(p.hub_info + 4 * h);
  const int first = __ldg(p.hub_info + 4 * h + 1), nch = __ldg(p.hub_info + 4 * h + 2);
  const int deg = __ldg(p.hub_info + 4 * h + 3);

  FeatMap<VEC, G, K> fm;
  fm.init(p, gl, blockIdx.y * (G * VEC * K));
  Acc<VEC> acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k].init();

  const bool two_level = nch > kFinGroups;
  if (two_level) {
    const int mine = (nch - q + kFinGroups - 1) / kFinGroups;     // chunks q, q+kFinGroups, ..
    merge_partials<VEC, K, UF>(p.partials, p.F, fm.f, fm.ok, first + q, mine, kFinGroups, acc);
    float* part = p.partials + (long long)(first + q) * 4ll * p.F;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!fm.ok[k]) continue;
      store_f32<VEC>(part + 0ll * p.F + fm.f[k], acc[k].sum);
      store_f32<VEC>(part + 1ll * p.F + fm.f[k], acc[k].sq);
      store_f32<VEC>(part + 2ll * p.F + fm.f[k], acc[k].mn);
      store_f32<VEC>(part + 3ll * p.F + fm.f[k], acc[k].mx);
    }
    __syncthreads();   // CTA-scope visibility of the parked per-group results
    if (q != 0) return;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k].init();
    merge_partials<VEC, K, UF>(p.partials, p.F, fm.f, fm.ok, first, kFinGroups, 1, acc);
  } else {
    if (q != 0) return;
    merge_partials<VEC, K, UF>(p.partials, p.F, fm.f, fm.ok, first, nch, 1, acc);
  }
  finalize_row<T, VEC, G, K, CfgDynamic>(p, fm, row, deg, acc);
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
      const unsigned std_s = (0u) | (1u << 4) | (2u << 8);
      const bool s3 = p.nS == 3 && (p.scodes & 0xfffu) == std_s && p.nA == 4;
      if (s3 && (p.acodes & 0xffffu) == CfgMeanMaxMinStd::ACODES)
        k_rows<T, VEC, G, K, U, CfgMeanMaxMinStd><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>(p);
      else if (s3 && (p.acodes & 0xffffu) == CfgMeanMinMaxStd::ACODES)
        k_rows<T, VEC, G, K, U, CfgMeanMinMaxStd><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>(p);
      else
        k_rows<T, VEC, G, K, U, CfgDynamic><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>(p);
      PNA_CUDA_TRY(cudaGetLastError());
    }
  }
  if (!(p.flags & PNA_FLAG_SKIP_HUBS) && p.n_hubs > 0) {
    const long long gc = (p.n_chunks + per_block - 1) / per_block;
    k_hub_chunks<T, VEC, G, K, U><<<dim3((unsigned)gc, gy), kThreads, 0, st>>>(p);
    PNA_CUDA_TRY(cudaGetLastError());
    k_hub_finalize<T, VEC, G, K><<<dim3((unsigned)p.n_hubs, gy), kFinGroups * G, 0, st>>>(p);
    PNA_CUDA_TRY(cudaGetLastError());
  }
  return PNA_OK;
}

template <typename T, int VEC>
int launch_typed(const KParams& p, cudaStream_t st) {
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

}  // namespace pna
