/*
 * Plain-C restatement of the PNA aggregation -- TEST INFRASTRUCTURE (second, independent oracle).
 *
 * Follows reference models/pytorch_geometric/pna.py:242-249 (aggregate), aggregators.py:9-32, scalers.py:8-29 with
 * torch_scatter's CPU semantics (sequential accumulation in edge order, empty segment -> 0).  Scalar fp32 loops, built
 * with -ffp-contract=off so no FMA is formed: it documents exactly which roundings the CUDA kernel reproduces.
 * Used by tests/ to cross-check oracle/pna_oracle.py (torch ops) and the accumulation-order claim; never by pna_b200/.
 *
 * out[i, (s*A + a)*F + f], aggr codes 0..5 = sum mean min max var std, scaler codes 0..4 = identity amplification
 * attenuation linear inverse_linear (same numbering as include/pna_b200.h).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int pna_oracle_aggregate(const float* x, int64_t n_nodes, int64_t n_feat, const int64_t* src, const int64_t* dst,
                         int64_t n_edges, const int32_t* aggr, int32_t n_aggr, const int32_t* scal, int32_t n_scal,
                         float avg_log, float avg_lin, int zero_isolated, float* out) {
  const int64_t N = n_nodes, F = n_feat, W = (int64_t)n_aggr * n_scal * F;
  float* sum = (float*)calloc((size_t)(N * F), sizeof(float));
  float* sq = (float*)calloc((size_t)(N * F), sizeof(float));
  float* mn = (float*)malloc((size_t)(N * F) * sizeof(float));
  float* mx = (float*)malloc((size_t)(N * F) * sizeof(float));
  int64_t* deg = (int64_t*)calloc((size_t)N, sizeof(int64_t));
  if (!sum || !sq || !mn || !mx || !deg) return -1;
  for (int64_t i = 0; i < N * F; ++i) { mn[i] = INFINITY; mx[i] = -INFINITY; }
  for (int64_t e = 0; e < n_edges; ++e) {           /* edge order == scatter_add_ order on CPU */
    const int64_t j = src[e], i = dst[e];
    if (j < 0 || j >= N || i < 0 || i >= N) return -2;
    deg[i]++;
    for (int64_t f = 0; f < F; ++f) {
      const float m = x[j * F + f];
      sum[i * F + f] = sum[i * F + f] + m;
      sq[i * F + f] = sq[i * F + f] + m * m;          /* src * src is rounded, then added (aggregators.py:27) */
      if (m < mn[i * F + f]) mn[i * F + f] = m;
      if (m > mx[i * F + f]) mx[i * F + f] = m;
    }
  }
  for (int64_t i = 0; i < N; ++i) {
    const int iso = deg[i] == 0;
    const float d = (float)deg[i];
    const float cnt = iso ? 1.0f : d;
    const float lg = logf(d + 1.0f);
    float scale[8];
    for (int s = 0; s < n_scal; ++s) {
      switch (scal[s]) {
        case 0: scale[s] = 1.0f; break;
        case 1: scale[s] = lg / avg_log; break;
        case 2: scale[s] = iso ? 1.0f : avg_log / lg; break;
        case 3: scale[s] = d / avg_lin; break;
        default: scale[s] = iso ? 1.0f : avg_lin / d; break;
      }
    }
    for (int64_t f = 0; f < F; ++f) {
      const float mean = sum[i * F + f] / cnt;
      const float msq = sq[i * F + f] / cnt;
      const float var = msq - mean * mean;
      const float sd = sqrtf((var > 0.0f ? var : 0.0f) + 1e-5f);
      for (int a = 0; a < n_aggr; ++a) {
        float v;
        switch (aggr[a]) {
          case 0: v = sum[i * F + f]; break;
          case 1: v = mean; break;
          case 2: v = iso ? 0.0f : mn[i * F + f]; break;
          case 3: v = iso ? 0.0f : mx[i * F + f]; break;
          case 4: v = var; break;
          default: v = sd; break;
        }
        for (int s = 0; s < n_scal; ++s)
          out[i * W + ((int64_t)s * n_aggr + a) * F + f] = (iso && zero_isolated) ? 0.0f : v * scale[s];
      }
    }
  }
  free(sum); free(sq); free(mn); free(mx); free(deg);
  return 0;
}
