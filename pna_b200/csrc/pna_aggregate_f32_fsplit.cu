// Feature-split instantiation of the streamed aggregation kernel: fp32 rows, 64-bit lane chunks (32 lanes x 2 features =
// a 64-feature block per pass), n_fpass passes over the warp's rows.  The gathered working set of one pass is
// n_src * 256 B instead of n_src * 512 B, so it stays L2-resident next to the output stream.
#include "pna_aggregate_impl.cuh"

namespace pna {

template <typename Cfg>
static int launch_fsplit(const KParams& p_in, cudaStream_t st) {
  constexpr int VEC = 2, K = 1, DEPTH = 1;
  KParams p = p_in;
  p.n_fpass = (p.F + 32 * VEC * K - 1) / (32 * VEC * K);
  p.work_ctr = nullptr;       // static assignment only on this experimental path
  constexpr size_t smem = StreamGeom<float, VEC, K, DEPTH>::kSmem;
  auto kern = k_rows_stream<float, VEC, K, Cfg, false, DEPTH, false>;
  static int resident = 0;
  if (resident == 0) {
    if (smem > 48 * 1024) PNA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, nb = 0;
    PNA_CUDA_TRY(cudaGetDevice(&dev));
    PNA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PNA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kStreamThreads, smem));
    resident = (nb > 0 ? nb : 1) * sms;
  }
  const long long slots = p.n_rows;
  long long gxs = (slots + 8 * (kStreamThreads / 32) - 1) / (8 * (kStreamThreads / 32));
  if (gxs > resident) gxs = resident;
  if (gxs < 1) gxs = 1;
  kern<<<dim3((unsigned)gxs, 1), kStreamThreads, smem, st>>>(p);
  PNA_CUDA_TRY(cudaGetLastError());
  return PNA_OK;
}

int launch_stream_fsplit_f32(const KParams& p, cudaStream_t st) {
  if (p.T != 1 || p.has_self || p.row_ids || (p.F % 64) != 0 || (p.ldx % 2) != 0 || (p.ldo % 2) != 0) return 1;
  return launch_fsplit<CfgMeanMaxMinStd>(p, st);
}

}  // namespace pna
