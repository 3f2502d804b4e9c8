// pna_aggregate_fwd: C-ABI entry point -- argument checks and dtype / alignment dispatch.
#include "pna_aggregate.cuh"

namespace pna {

template <typename T, int VEC>
int launch_typed(const KParams& p, cudaStream_t st);   // defined in pna_aggregate_{f32,bf16}_{vec,scalar}.cu

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
    PNA_REQUIRE(((d->aggr_codes >> (4 * a)) & 15u) <= PNA_AGGR_STD || ((d->aggr_codes >> (4 * a)) & 15u) == PNA_AGGR_SKIP,
                PNA_ERR_BAD_ARG, "pna_aggregate_fwd: bad aggregator code");
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
  p.hub_done = d->hub_done;
  p.row_ids = d->row_ids; p.n_row_ids = d->row_ids ? d->n_row_ids : 0;
  // a view that contains chunk pseudo-rows cannot be used when the split rows are to be skipped
  const bool view = d->light_rowptr && d->light_deg && d->part && d->n_part >= 1 && (d->light_col || !d->col) &&
                    !(d->n_view_rows > d->n_rows && ((d->flags & PNA_FLAG_SKIP_HUBS) || d->row_ids));
  p.lrowptr = view ? d->light_rowptr : nullptr; p.ldeg = view ? d->light_deg : nullptr; p.lcol = d->light_col; p.part = d->part;
  p.n_part = d->n_part;
  // chunk pseudo-rows are only usable when the split rows are to be processed in this call
  p.n_view_rows = (view && d->n_view_rows > d->n_rows && !(d->flags & PNA_FLAG_SKIP_HUBS) && d->n_hubs > 0 && !d->row_ids)
                      ? d->n_view_rows : d->n_rows;
  PNA_REQUIRE(p.n_view_rows == d->n_rows || d->n_view_rows == d->n_rows + d->n_chunks, PNA_ERR_BAD_ARG,
              "pna_aggregate_fwd: n_view_rows must be n_rows + n_chunks");
  p.n_fpass = 0;
  p.work_ctr = d->work_counter; p.n_static = 0;
  p.sdeg = d->scaler_degree;
  // more than 512 chunks in one row: merge the partials with the radix tree (k_hub_tree) before the finalize
  p.hub_merged = (d->max_degree > 0 && (long long)d->max_degree > 512ll * d->chunk_edges) ? 1 : 0;
  p.peer_x = reinterpret_cast<const void* const*>(d->peer_gathered); p.peer_shift = d->peer_shift;
  if (p.peer_x) PNA_REQUIRE(d->peer_shift >= 1 && d->peer_shift <= 30, PNA_ERR_BAD_ARG, "pna_aggregate_fwd: peer_shift out of range");
  PNA_REQUIRE(p.ldx < 0x3fffffffll && p.ldb < 0x7fffffffll && p.lds < 0x7fffffffll, PNA_ERR_UNSUPPORTED,
              "pna_aggregate_fwd: row pitch too large");     // ld_gathered in BYTES is a 32-bit kernel operand
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
