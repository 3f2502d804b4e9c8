/*
 * pna_b200.h -- C ABI of libpna_sm100.so, the B200 (sm_100a) PNA aggregation path.
 *
 * This is the drop-in boundary for ONE hot path of lukecavabarrett/pna: the
 * neighbourhood aggregation of the PNA layer (gather of source features, the
 * simultaneous mean/max/min/std(/sum/var) aggregators, the degree scalers, the
 * concatenated [N, S*A*F] result).  Plain pointers and sizes only; every buffer
 * is owned by the caller (PyTorch in this repo); all pointers are DEVICE pointers
 * unless stated otherwise; every call enqueues on the caller's stream.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   models/pytorch_geometric/pna.py:152-159, :242-249   PNAConv(.Simple).aggregate
 *   models/pytorch_geometric/aggregators.py:9-32        scatter sum/mean/min/max/var/std
 *   models/pytorch_geometric/scalers.py:8-29            identity/amplification/attenuation/linear/inverse_linear
 *   models/dgl/pna_layer.py:45-50, :189-194             PNATower.reduce_func / PNASimpleLayer.reduce_func
 *   models/dgl/aggregators.py:6-26, models/dgl/scalers.py:8-19
 *   torch_geometric MessagePassing.propagate gather of x_j (pna.py:129, :236)
 *
 * Return value of every int function: 0 = ok, negative = pna_status.
 * pna_last_error() gives a thread-local human-readable message for the last failure.
 */
#ifndef PNA_B200_H
#define PNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNA_ABI_VERSION 8

typedef void* pna_stream_t; /* a cudaStream_t / CUstream, passed opaquely */

enum pna_status {
  PNA_OK = 0,
  PNA_ERR_BAD_ARG = -1,     /* null pointer, negative size, inconsistent descriptor */
  PNA_ERR_UNSUPPORTED = -2, /* dtype / size outside what the kernels take (e.g. E >= 2^31) */
  PNA_ERR_CUDA = -3,        /* a CUDA runtime call failed; message carries cudaGetErrorString */
  PNA_ERR_INDEX = -4,       /* an edge endpoint outside [0, n_nodes) was found while building the CSR */
  PNA_ERR_WORKSPACE = -5    /* caller-provided workspace / capacity too small */
};

enum pna_dtype { PNA_F32 = 0, PNA_BF16 = 1 };

/* Aggregator codes (order of appearance in the layer's ctor list fixes the column layout,
 * pna.py:70,153-154).  Packed 4 bits each, first aggregator in the low nibble. */
enum pna_aggr { PNA_AGGR_SUM = 0, PNA_AGGR_MEAN = 1, PNA_AGGR_MIN = 2, PNA_AGGR_MAX = 3, PNA_AGGR_VAR = 4, PNA_AGGR_STD = 5,
                PNA_AGGR_SKIP = 15 /* keep the column slot but do not write it: lets two calls with different edge
                                       sets / messages fill one output row (dense reference layer, where max/min and
                                       mean/std see different messages: models/pytorch/pna/aggregators.py:30-51) */ };
/* Scaler codes (scalers.py:32-38), packed the same way. */
enum pna_scaler { PNA_SCALE_IDENTITY = 0, PNA_SCALE_AMPLIFICATION = 1, PNA_SCALE_ATTENUATION = 2, PNA_SCALE_LINEAR = 3, PNA_SCALE_INVERSE_LINEAR = 4 };

#define PNA_MAX_AGGR 6
#define PNA_MAX_SCALERS 5

enum pna_flags {
  PNA_FLAG_ZERO_ISOLATED = 1u, /* DGL semantics: a node with no in-edge is never reduced, all its S*A*F columns are 0
                                  (models/dgl/pna_layer.py:64 update_all).  Default (PyG/torch_scatter semantics):
                                  mean=min=max=0, std=sqrt(1e-5), then scaled. */
  PNA_FLAG_SKIP_LIGHT = 2u,    /* do not process rows below the split threshold (used to overlap halo exchange) */
  PNA_FLAG_SKIP_HUBS = 4u,     /* do not process rows at/above the split threshold */
  PNA_FLAG_GATHER_L1 = 16u,    /* hint: a few source rows receive a large share of all gathers (power-law graphs).  The
                                  gathered rows are then also kept in the SMs' L1, so a hot row is not served by the few L2
                                  slices that hold its lines.  Without hot sources the hint costs a few percent.  Honoured by
                                  the streamed kernel for the reference configs' aggregator / scaler lists. */
  PNA_FLAG_RELU_VAR = 8u       /* the "var" aggregator is clamped at 0 (models/dgl/aggregators.py:22-26 and
                                  models/pytorch/pna/aggregators.py:63-76 apply torch.relu; the PyG flavour,
                                  models/pytorch_geometric/aggregators.py:25-28, does not); its gradient is masked where
                                  the raw variance is <= 0 */
};

enum pna_query_what {
  PNA_QUERY_ABI_VERSION = 0,
  PNA_QUERY_SM_ARCH = 1,          /* 100 : compiled for sm_100a only */
  PNA_QUERY_DEFAULT_SPLIT = 2,    /* default in-degree at which a row is split across warps */
  PNA_QUERY_DEFAULT_CHUNK = 3,    /* default edges per chunk of a split row */
  PNA_QUERY_DEVICE_SM_COUNT = 4,  /* multiProcessorCount of the current device (needs a GPU) */
  PNA_QUERY_MAX_FEATURES = 5,     /* largest n_feat accepted by pna_aggregate_fwd */
  PNA_QUERY_SIZEOF_CSR = 6,       /* sizeof(pna_csr_t): lets an FFI binding verify its struct layout */
  PNA_QUERY_SIZEOF_AGG = 7        /* sizeof(pna_agg_t) */
};

/* ---- destination-sorted CSR ("sorts/segments edges by destination in CSR", north_star) ------------------
 * Replaces what torch_scatter does implicitly on every call (scatter by edge_index[1], pna.py:153,157).
 * Edges are ordered by destination, ties kept in original edge order (stable), duplicates and self loops kept:
 *   rowptr[i]..rowptr[i+1]  = CSR slots of the in-edges of node i;  in-degree = difference
 *   col[s]                  = source node of slot s
 *   perm[s]                 = original edge id of slot s (to bring per-edge tensors into CSR order)
 * Rows with in-degree >= split_threshold ("hubs") are additionally listed with a chunking of their slots so the
 * aggregation can spread them over many warps:
 *   hub_info[4*h+0..3]      = row, first chunk id, number of chunks, in-degree
 *   chunk_items[2*c+0..1]   = hub index h, chunk index within the hub
 */
typedef struct pna_csr {
  int64_t n_nodes;          /* in */
  int64_t n_edges;          /* in */
  int32_t split_threshold;  /* in: >= 2 */
  int32_t chunk_edges;      /* in: 1..split_threshold */
  int32_t* rowptr;          /* out [n_nodes+1] */
  int32_t* col;             /* out [n_edges] */
  int32_t* perm;            /* out [n_edges] */
  int32_t* hub_info;        /* out [4*cap_hubs] */
  int32_t* chunk_items;     /* out [2*cap_chunks] */
  int64_t cap_hubs;         /* in: >= n_edges/split_threshold + 1 */
  int64_t cap_chunks;       /* in: >= n_edges/chunk_edges + cap_hubs + 1 (and <= 2*n_edges + 3 when the light view is requested) */
  int64_t n_hubs;           /* out (host) */
  int64_t n_chunks;         /* out (host) */
  int32_t max_degree;       /* out (host) */
  int32_t n_part;           /* in: number of equal-cost row partitions to produce (>= 1) */
  /* "light view": the slots of the rows BELOW the split threshold, compacted so that any contiguous range of rows
   * is a contiguous range of slots -- what lets the streaming kernel treat a warp's rows as one slot stream.
   * Optional: pass light_rowptr == NULL to skip it. */
  /* The view built here has n_nodes + cap_chunks rows: after the real rows comes one PSEUDO-ROW per chunk of the
   * split rows (n_chunks valid), whose slots the same kernel reduces into hub_partials -- so all gathers of a layer
   * call run in one balanced launch. */
  int32_t* light_rowptr;    /* out [n_nodes+cap_chunks+1]: prefix sum of the view rows' slot counts (split rows: 0) */
  int32_t* light_deg;       /* out [n_nodes+cap_chunks]: slot count of each view row, -1 = skip (split row, unused chunk row) */
  int32_t* light_col;       /* out [n_edges]: source node of each view slot (first n_light_edges entries valid) */
  int32_t* part;            /* out [n_part+1]: view-row boundaries of partitions of equal cost (slots + 12 * rows) */
  int64_t n_light_edges;    /* out (host): slots in the view (= n_edges when chunk rows are included) */
  int64_t n_src_nodes;      /* in: sources are validated against [0, n_src_nodes); 0 = n_nodes.  > n_nodes for the
                               destination-partitioned multi-GPU path, where sources index [local rows ; halo rows] */
  float hot_source_fraction; /* out (host): estimated share of the gathers that go to frequent source rows (sampled: sources
                                seen >= 4 times among 65 536 evenly spaced slots).  ~0 for citation / molecule / kNN graphs,
                                ~0.9 for Zipf-distributed sources; above ~0.25 pass PNA_FLAG_GATHER_L1 to the aggregation */
  int32_t reserved;
} pna_csr_t;

/* Bytes of device scratch pna_csr_build needs for (n_nodes, n_edges) on the current device. */
int pna_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, size_t* bytes);

/* src[e] -> dst[e], e in [0, n_edges): int64 device arrays (edge_index[0], edge_index[1] of PyG;
 * g.edges() of DGL).  Synchronises the stream once at the end to return the host-side counts and to report
 * out-of-range endpoints (PNA_ERR_INDEX).  Call once per graph, not per layer. */
int pna_csr_build(const int64_t* src, const int64_t* dst, pna_csr_t* csr, void* workspace, size_t workspace_bytes,
                  pna_stream_t stream);

/* Light view restricted to the rows with row_mask[r] != 0 (NULL = all rows below the split threshold); rows outside
 * the mask get light_deg = -1 and no slots, so the streaming kernel skips them without a row list.  Used to run the
 * rows whose sources are all local while the halo all-to-all is in flight, then the rest.  workspace: at least
 * pna_csr_light_view_workspace_bytes(n_nodes) bytes of device scratch. */
int pna_csr_light_view(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, int32_t split_threshold, const uint8_t* row_mask,
                       int32_t n_part, int32_t* light_rowptr, int32_t* light_deg, int32_t* light_col, int32_t* part,
                       void* workspace, size_t workspace_bytes, pna_stream_t stream);
int pna_csr_light_view_workspace_bytes(int64_t n_nodes, size_t* bytes);

/* ---- the aggregation ("single hand-written sm_100a CUDA kernel", north_star) ------------------------------
 * For every destination row i (PyG semantics; In(i) = slots rowptr[i]..rowptr[i+1], d = |In(i)|):
 *   m_s   = gathered[col[s]] (+ row_bias[i] when given)          s in In(i)           pna.py:137-150 / :239-240
 *   sum   = fp32 sum of m_s in slot order;  mean = sum / max(d,1)                      aggregators.py:9-14
 *   min/max over m_s, 0 when d == 0                                                     aggregators.py:17-22
 *   var   = (sum of m_s*m_s)/max(d,1) - mean*mean ; std = sqrt(max(var,0) + 1e-5)       aggregators.py:25-32
 *   amplification = log(d+1)/avg_log ; attenuation = d ? avg_log/log(d+1) : 1           scalers.py:12-19
 *   linear = d/avg_lin ; inverse_linear = d ? avg_lin/d : 1                             scalers.py:22-29
 *   out[i, tower t, ((s*A + a)*Ft + f)] = scaler_s * aggr_a                            pna.py:154-159 (scaler-major)
 * With n_towers = T the n_feat columns are T blocks of Ft = n_feat/T, and the output row is T blocks of
 * (has_self + S*A)*Ft columns, i.e. the [N, T, (1+)S*A*Ft] tensor of pna.py:129-131 flattened; when self_feat is
 * given, block t starts with self_feat[i, t*self_tower_stride : +Ft] (the torch.cat([x, out]) of pna.py:131).
 * All accumulation, degree, log and scaling in fp32 for both dtypes; bf16 is converted on load / store.
 */
typedef struct pna_agg {
  const void* gathered;      /* [n_src, n_feat] rows to gather (x for PNAConvSimple, V = x W_j^T + b for PNAConv) */
  int64_t ld_gathered;       /* row pitch in elements */
  const int32_t* rowptr;     /* [n_rows+1] */
  const int32_t* col;        /* [n_edges]; NULL = identity (gathered[] holds per-edge messages already in CSR order) */
  const void* row_bias;      /* nullable [n_rows, n_feat]: destination-side term added to every gathered row */
  int64_t ld_row_bias;
  const void* self_feat;     /* nullable: prepend self features to every tower block of the output row */
  int64_t ld_self;
  int64_t self_tower_stride; /* Ft when the input is divided between towers, 0 when it is repeated (pna.py:123-126) */
  void* out;                 /* [n_rows, ld_out] */
  int64_t ld_out;            /* >= n_towers * (has_self + S*A) * Ft */
  int64_t n_rows;
  int32_t n_feat;            /* total gathered width = n_towers * Ft */
  int32_t n_towers;
  int32_t dtype;             /* pna_dtype of gathered / row_bias / self_feat / out */
  int32_t n_aggr;
  uint32_t aggr_codes;
  int32_t n_scalers;
  uint32_t scaler_codes;
  float avg_log;             /* avg_deg['log'] of the layer ctor (pna.py:84) */
  float avg_lin;             /* avg_deg['lin'] (pna.py:83); only read by linear / inverse_linear */
  uint32_t flags;            /* pna_flags */
  int32_t split_threshold;   /* must equal the value the CSR was built with */
  int32_t chunk_edges;
  const int32_t* hub_info;   /* from pna_csr_t; may be NULL when n_hubs == 0 */
  const int32_t* chunk_items;
  int64_t n_hubs;
  int64_t n_chunks;
  float* hub_partials;       /* fp32 scratch [n_chunks * 4 * n_feat] (pna_aggregate_bwd: [(n_chunks + n_hubs) * 6 * n_feat]);
                                may be NULL when n_hubs == 0 */
  const int32_t* row_ids;    /* nullable [n_row_ids]: process only these light rows (halo overlap); hubs unaffected.
                                With a light view, view row i is output row row_ids[i]. */
  int64_t n_row_ids;
  /* optional light view of the rows (pna_csr_t light_* / part; NULL = not available -> tile kernels on rowptr/col).
   * Without row_ids it has n_rows rows; with row_ids it has n_row_ids rows. */
  const int32_t* light_rowptr;
  const int32_t* light_deg;
  const int32_t* light_col;
  const int32_t* part;
  int32_t n_part;
  /* view rows beyond n_rows are chunk pseudo-rows (row n_rows + c reduces chunk c into hub_partials); 0 or n_rows = none */
  int64_t n_view_rows;
  /* destination-partitioned multi-GPU graph, gather fused with the exchange: peer_gathered[r] (DEVICE array of
   * n_ranks device pointers) is rank r's `gathered` buffer mapped into this process (CUDA IPC / symmetric memory over
   * NVLink); a col entry c then means row (c & ((1 << peer_shift) - 1)) of rank (c >> peer_shift).  NULL: single GPU,
   * col indexes `gathered` directly.  All ranks use the same ld_gathered. */
  const void* const* peer_gathered;
  int32_t peer_shift;
  int32_t max_degree;        /* pna_csr_t.max_degree, or 0 = unknown.  A row with more than 64 * 8 chunks (power-law graphs:
                              * millions of in-edges) has its chunk partials merged by a radix tree of small launches over the
                              * chunk array instead of by one CTA walking them -- same fixed order on every call */
  /* optional int32 [9 * n_hubs] completion counters, ZERO before the first call (the library leaves them zero): when
   * given together with a view that contains the chunk pseudo-rows, the warp that stores the last partial of a split row
   * also merges and finalizes it (same merge order as the separate finalize kernel), so a layer call is ONE launch with
   * no serial tail.  NULL: split rows are finalized by a second small kernel.  Not to be shared by concurrent calls. */
  int32_t* hub_done;
  /* optional int32 [n_rows]: the degree the SCALERS see, when it is not the in-degree of the CSR row.  The dense reference
   * layer aggregates over adj + I with self_loop=True but scales with D = adj.sum(-1) of the loop-free adjacency
   * (models/pytorch/pna/scalers.py:13,21,28,35), and its max/min reduce over the other adjacency axis while still being
   * scaled with the row degree.  NULL: scalers use rowptr[i+1] - rowptr[i] (PyG / DGL). */
  const int32_t* scaler_degree;
  /* optional: ONE int32 of device scratch (the library zeroes it on the stream before the launch).  When given, the streamed
   * kernel deals out only the first 70 % of the row partitions statically and hands out the rest one at a time through
   * this counter, so warps that finish their static range early take over work from slow ones.  Not to be shared by
   * concurrent calls.  NULL: fully static assignment. */
  int32_t* work_counter;
} pna_agg_t;

int pna_aggregate_fwd(const pna_agg_t* desc, pna_stream_t stream);

/* Backward of pna_aggregate_fwd w.r.t. the gathered rows (needed by every training loop,
 * multitask_benchmark/util/train.py:148).  grad_out has the layout of out (self block, if any, is skipped);
 * grad_gathered [n_src, n_feat] fp32 must be zero-initialised by the caller, contributions are accumulated with
 * atomics; grad_row_bias (nullable) [n_rows, n_feat] fp32 is written.  min/max route to the first slot attaining
 * the extremum (torch_scatter arg semantics). */
int pna_aggregate_bwd(const pna_agg_t* desc, const void* grad_out, int64_t ld_grad_out, float* grad_gathered,
                      int64_t ld_grad_gathered, float* grad_row_bias, int64_t ld_grad_row_bias, pna_stream_t stream);

/* The same gradient without one atomic per (edge, feature), for gathered rows (desc->col != NULL) -- three calls:
 *  1. pna_aggregate_bwd_coef: per destination row i the gradient of a message is  c0_i + c1_i * m + routed min / max terms;
 *     writes coef[i] = [c0_i + c1_i * row_bias[i]  (n_feat floats at column 0) | c1_i (n_feat floats at column c1_column)]
 *     for every row with in-edges (other rows are left untouched and are never read in step 2), adds the min / max
 *     gradients to grad_gathered[col[arg slot]] -- one scalar atomic per (row, feature); grad_gathered zero-initialised by
 *     the caller as above -- and writes grad_row_bias (nullable).  c1_column >= n_feat, ld_coef >= c1_column + n_feat;
 *     16-byte aligned choices (c1_column % 4 == 0, ld_coef % 4 == 0) get vector stores.  desc->hub_partials as for
 *     pna_aggregate_bwd.
 *  2. the caller sums the coefficient rows over the out-edges of every source row: pna_aggregate_fwd on the CSR of the
 *     TRANSPOSED graph (pna_csr_build with source and destination swapped: n_nodes = n_src) with gathered = coef,
 *     n_feat = ld_coef, one aggregator PNA_AGGR_SUM, one scaler PNA_SCALE_IDENTITY -> sums [n_src, ld_coef].
 *  3. pna_aggregate_bwd_combine: grad_gathered[j, f] += sums[j, f] + gathered[j, f] * sums[j, c1_column + f].
 * Same results up to fp32 summation order (reference: autograd of aggregators.py:9-32; tests/test_bwd_two_phase_math.py
 * restates the algebra). */
int pna_aggregate_bwd_coef(const pna_agg_t* desc, const void* grad_out, int64_t ld_grad_out, float* coef, int64_t ld_coef,
                           int32_t c1_column, float* grad_gathered, int64_t ld_grad_gathered, float* grad_row_bias,
                           int64_t ld_grad_row_bias, pna_stream_t stream);
int pna_aggregate_bwd_combine(const float* coef_sums, int64_t ld_sums, int32_t c1_column, const void* gathered,
                              int64_t ld_gathered, int32_t dtype, float* grad_gathered, int64_t ld_grad_gathered, int64_t n_src,
                              int32_t n_feat, pna_stream_t stream);

/* ---- halo rows for the destination-partitioned multi-GPU path (north_star: "single NCCL all-to-all for halo
 * source features per layer"): dst[i, :] = src[idx[i], :], n_feat elements per row.  Used to pack the send buffer. */
int pna_gather_rows(const void* src, int64_t ld_src, const int32_t* idx, int64_t n_idx, void* dst, int64_t ld_dst,
                    int32_t n_feat, int32_t dtype, pna_stream_t stream);

/* ---- the same exchange as ONE kernel of peer loads (no pack, no collective, no unpack) -------------------------------
 * Every rank's feature rows live in a buffer that is mapped into every process (symmetric memory / CUDA IPC over NVLink).
 * peer_rows: DEVICE array of n_ranks base pointers of those buffers (row pitch ld_rows elements on every rank);
 * enc[i] = owner << peer_shift | row-on-owner names the i-th de-duplicated remote source row this rank needs;
 * dst[i, :] (pitch ld_dst; normally the tail of the rank's [local ; halo] buffer) receives it.  One remote row crosses
 * NVLink once per layer however many local destinations gather it afterwards. */
int pna_halo_pull(const void* const* peer_rows, int64_t ld_rows, const int32_t* enc, int32_t peer_shift, int64_t n_idx, void* dst,
                  int64_t ld_dst, int32_t n_feat, int32_t dtype, pna_stream_t stream);

/* Device-side barrier between the ranks of one box, enqueued on `stream`: peer_flags is a DEVICE array of n_ranks pointers to
 * each rank's uint64 flags[n_ranks] (zero-initialised once, mapped into every process like the feature rows).  Rank r
 * stores `epoch` into flags[r] of every peer and waits until its own flags all reach `epoch`; epochs must increase by
 * one per call on every rank.  Orders "every rank has written its feature rows" before the pulls / peer gathers of the
 * next kernel.  If a peer does not arrive within timeout_ns (0 = 2 s) the kernel gives up and sets *status = 1 (DEVICE int,
 * nullable) instead of hanging the GPU. */
int pna_peer_barrier(const void* const* peer_flags, int32_t rank, int32_t world, uint64_t epoch, uint64_t timeout_ns,
                     int32_t* status, pna_stream_t stream);

/* ---- first dense linear of the post-aggregation MLP on the tensor cores (north_star: "the post-MLP uses tensor
 * cores only for its dense linear"; reference pna.py:222-227 post_nn[0], models/dgl/pna_layer.py:31 posttrans) -----
 * y[n_rows, n_out] = a[n_rows, n_in] . weight[n_out, n_in]^T + bias, fp32 in / fp32 out, fp32-accurate: every operand
 * is split hi + lo and three tcgen05.mma kind::tf32 products (hi.hi + hi.lo + lo.hi) accumulate in TMEM, because plain
 * TF32 (10-bit mantissa) cannot meet the 1e-5 parity bar.  n_in % 32 == 0, n_out in {64, 128, 256}; other shapes
 * return PNA_ERR_UNSUPPORTED and the caller keeps its library GEMM.  workspace: 2 * n_in * n_out floats (split weight). */
int pna_linear_workspace_bytes(int32_t n_in, int32_t n_out, size_t* bytes);
int pna_linear_fwd(const float* a, int64_t lda, const float* weight, const float* bias, float* y, int64_t ldy, int64_t n_rows,
                   int32_t n_in, int32_t n_out, void* workspace, size_t workspace_bytes, pna_stream_t stream);

/* ---- the same linear fed by the COMPACT aggregate (the [N, S*A*F] tensor is never written) --
 * The reference's post-MLP input is cat over the scalers s of  scale_s(deg_i) * agg_i  (pna.py:247-249,
 * scalers.py:8-29): S scaled copies of one [N, A*F] tensor.  With a = that tensor for the identity scaler alone
 * (pna_aggregate_fwd with n_scalers = 1, scaler_codes = PNA_SCALE_IDENTITY) and row_scale[i, s] = the factor of scaler s
 * for row i (pna_row_scales),
 *     y[i, :] = sum_s sum_k  fl(row_scale[i, s] * a[i, k]) * weight[:, s * n_a + k]  + bias,      n_a = n_in / n_scalers
 * which is the reference's  post_nn[0](cat_s(...))  with every product rounded as the reference rounds it; the scaled
 * copies exist only in registers.  weight keeps the reference's layout [n_out, n_in = S*A*F] (scaler-major columns).
 * n_a % 32 == 0, n_out in {64, 128, 256}; workspace as for pna_linear_fwd. */
int pna_linear_scaled_fwd(const float* a, int64_t lda, const float* row_scale, int32_t n_scalers, const float* weight,
                          const float* bias, float* y, int64_t ldy, int64_t n_rows, int32_t n_in, int32_t n_out, void* workspace,
                          size_t workspace_bytes, pna_stream_t stream);
/* scales[i, s] = factor of scaler s (code (scaler_codes >> 4s) & 15) at in-degree rowptr[i+1] - rowptr[i]; bit-identical
 * to the factors pna_aggregate_fwd applies (one device function computes both). */
int pna_row_scales(const int32_t* rowptr, int64_t n_rows, int32_t n_scalers, uint32_t scaler_codes, float avg_log, float avg_lin,
                   float* scales, pna_stream_t stream);

int pna_query(int what);
const char* pna_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PNA_B200_H */
