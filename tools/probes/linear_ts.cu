// Round-2 candidate for pna_linear(_scaled)_fwd: the same 3xTF32 GEMM with the A operand
// staged in TENSOR MEMORY instead of shared memory.  Stand-alone experiment: NOT part of libpna_sm100.so, not run by the
// tests; run tools/probes/probe_tmem_a.cu first (it checks the TMEM cell <-> A[m][k] hypothesis this kernel relies on).
//
//   TMEM columns (512):  [0,128) main acc 0 | [128,256) main acc 1 | [256,384) cross-term acc | [384,512) A ring:
//                        2 stages x (32 columns A_hi | 32 columns A_lo), lane = row of the 128-row tile
//   warps 0-3 / 4-7      A loaders, thread = row: 8 x LDG.128 of the row's 32-wide K block (next block prefetched in
//                        registers), optional row scale (compact operand), cvt.rna.tf32 hi/lo split, 2 x tcgen05.st.x32.
//                        Group g = warp / 4 owns the K blocks with kb % 2 == g (all of their scaler copies).
//                        Warps 0-3 are afterwards the epilogue.
//   warp 8               MMA issuer: 12 tcgen05.mma per step, A from TMEM (TS form), W from shared memory
//   warp 9               W producer: bulk copies of the pre-swizzled weight images into a 6-deep ring (A frees 96 KB)
//
// Build + run (GPU box), compares against the library kernel and times both:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o /tmp/linear_ts tools/probes/linear_ts.cu && /tmp/linear_ts
#include "../../pna_b200/csrc/pna_misc.cu"
#include "../../pna_b200/csrc/pna_linear.cu"
#include <vector>
#include <random>
#include <cmath>
#include <functional>
#include <algorithm>

namespace pna {

constexpr int kTsO = 128;
constexpr int kTsWStages = 6;
constexpr int kTsWTile = kTsO * 128;               // bytes of one W hi (or lo) tile
constexpr int kTsWStage = 2 * kTsWTile;            // hi + lo
constexpr size_t kTsSmem = 1024 + (size_t)kTsWStages * kTsWStage + 256;
constexpr unsigned kTsACol = 384;                  // first TMEM column of the A ring
constexpr int kTsThreads = 320;

__device__ __forceinline__ void ts_mma(unsigned d, unsigned a_tmem, unsigned long long bdesc, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ts_st32(unsigned taddr, const unsigned (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

// barriers (8 bytes each, after the W ring): wfull[SW] wempty[SW] afull[2] aempty[2] done, then the TMEM slot
__global__ void __launch_bounds__(kTsThreads, 1)
k_linear_ts(const float* __restrict__ A, long long lda, const float* __restrict__ row_scale, int n_rep, const float* __restrict__ Wimg,
            const float* __restrict__ bias, float* __restrict__ Y, long long ldy, long long N, int K) {
  constexpr int O = kTsO, SW = kTsWStages;
  extern __shared__ unsigned char ts_raw[];
  const unsigned base = (lin_smem_u32(ts_raw) + 1023u) & ~1023u;
  unsigned char* gbase = ts_raw + (base - lin_smem_u32(ts_raw));
  const unsigned bars = base + SW * kTsWStage;
  const unsigned wfull = bars, wempty = bars + 8 * SW, afull = bars + 16 * SW, aempty = afull + 16, done = aempty + 16, slot = done + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * kLinM;
  const int n_kb = K / kLinBK, n_it = n_kb * n_rep;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SW; ++s) { lin_mbar_init(wfull + 8 * s, 1); lin_mbar_init(wempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { lin_mbar_init(afull + 8 * s, 4); lin_mbar_init(aempty + 8 * s, 1); }   // 4 loader warps per step
    lin_mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *reinterpret_cast<volatile unsigned*>(gbase + (slot - base));

  if (warp < 8) {
    // ---------------- A loaders: thread = row ----------------
    const int grp = warp >> 2, q = warp & 3;                  // K-block parity owned / TMEM lane quarter
    const long long r = row0 + q * 32 + lane;
    const bool live = r < N;
    const float* arow = A + (live ? r : 0) * lda;
    float4 cur[8], nxt[8];
    auto fetch = [&](int kb, float4 (&dst)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = (live && kb < n_kb) ? __ldg(reinterpret_cast<const float4*>(arow + kb * kLinBK + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch(grp, nxt);
    for (int kb = grp; kb < n_kb; kb += 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      fetch(kb + 2, nxt);                                     // in flight while this block's copies are produced
      for (int rep = 0; rep < n_rep; ++rep) {
        const int it = kb * n_rep + rep;                      // pipeline step
        const int sa = it & 1;
        const unsigned use = (unsigned)(it >> 1);             // how many times this A stage has been used before
        const float sc = row_scale ? (live ? __ldg(row_scale + r * n_rep + rep) : 0.f) : 1.f;
        lin_mbar_wait(aempty + 8 * sa, (use & 1) ^ 1);        // MMAs of the previous use have read the stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + kTsACol + (unsigned)sa * 64u;
        unsigned part[32];                                    // hi first, then lo: one 32-register staging array
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x = row_scale ? __fmul_rn(v[e], sc) : v[e];
              const float h = lin_tf32(x);
              part[j * 4 + e] = __float_as_uint(half == 0 ? h : lin_tf32(x - h));
            }
          }
          ts_st32(taddr + 32u * half, part);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) lin_mbar_arrive(afull + 8 * sa);
      }
    }
    if (warp < 4) {
      // ---------------- epilogue (as in k_linear_3xtf32) ----------------
      lin_mbar_wait(done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long row = row0 + warp * 32 + lane;
#pragma unroll
      for (int c0 = 0; c0 < O; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (a == 1 && n_it <= 1) continue;
          unsigned c[16];
          const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(a * O + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]),
                "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = (a == 0) ? __uint_as_float(c[i]) : acc[i] + __uint_as_float(c[i]);
        }
        if (row < N) {
          float* yr = Y + row * ldy + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o;
            o.x = acc[i] + (bias ? __ldg(bias + c0 + i) : 0.f);
            o.y = acc[i + 1] + (bias ? __ldg(bias + c0 + i + 1) : 0.f);
            o.z = acc[i + 2] + (bias ? __ldg(bias + c0 + i + 2) : 0.f);
            o.w = acc[i + 3] + (bias ? __ldg(bias + c0 + i + 3) : 0.f);
            *reinterpret_cast<float4*>(yr + i) = o;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(O >> 3) << 17) | ((unsigned)(kLinM >> 4) << 24);
    for (int it = 0; it < n_it; ++it) {
      const int sa = it & 1, sw = it % SW;
      lin_mbar_wait(afull + 8 * sa, (unsigned)(it >> 1) & 1u);
      lin_mbar_wait(wfull + 8 * sw, (unsigned)(it / SW) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const unsigned w_hi = base + sw * kTsWStage, w_lo = w_hi + kTsWTile;
        const unsigned a_hi = tmem + kTsACol + (unsigned)sa * 64u, a_lo = a_hi + 32u;
        const unsigned corr = tmem + 2 * O, mainacc = tmem + (unsigned)(it & 1) * O;
#pragma unroll
        for (int ks = 0; ks < kLinBK / 8; ++ks) {
          const unsigned ko = ks * 32, kc = ks * 8;
          ts_mma(corr, a_hi + kc, lin_desc(w_lo + ko), idesc, (it | ks) ? 1u : 0u);
          ts_mma(corr, a_lo + kc, lin_desc(w_hi + ko), idesc, 1u);
          ts_mma(mainacc, a_hi + kc, lin_desc(w_hi + ko), idesc, (it >= 2 || ks) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(aempty + 8 * sa) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wempty + 8 * sw) : "memory");
        if (it == n_it - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
      }
      __syncwarp();
    }
  } else if (warp == 9 && lane == 0) {
    // ---------------- W producer ----------------
    for (int it = 0; it < n_it; ++it) {
      const int sw = it % SW;
      const int kbw = (it % n_rep) * n_kb + it / n_rep;
      lin_mbar_wait(wempty + 8 * sw, ((unsigned)(it / SW) & 1u) ^ 1u);
      const unsigned st = base + sw * kTsWStage;
      lin_mbar_expect_tx(wfull + 8 * sw, 2u * kTsWTile);
      const float* img = Wimg + (long long)kbw * (2 * O * kLinBK);
      lin_bulk_g2s(st, img, kTsWTile, wfull + 8 * sw);
      lin_bulk_g2s(st + kTsWTile, img + O * kLinBK, kTsWTile, wfull + 8 * sw);
    }
  }
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace pna

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static float time_ms(cudaStream_t st, int iters, const std::function<void()>& fn) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) fn();
  CK(cudaEventRecord(a, st));
  for (int i = 0; i < iters; ++i) fn();
  CK(cudaEventRecord(b, st));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / iters;
}

int main(int argc, char** argv) {
  const long long N = argc > 1 ? atoll(argv[1]) : 169343;
  const int KA = 512, S = 3, O = 128, K = KA * S;
  std::mt19937 rng(1);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(0.f, 3.f);
  std::vector<float> hA((size_t)N * KA), hC((size_t)N * S), hW((size_t)O * K), hB(O);
  for (auto& v : hA) v = nd(rng);
  for (size_t i = 0; i < hC.size(); ++i) hC[i] = (i % S == 0) ? 1.f : ud(rng);
  for (auto& v : hW) v = nd(rng) / std::sqrt((float)K);
  for (auto& v : hB) v = nd(rng);
  float *dA, *dC, *dW, *dB, *dY0, *dY1, *dWs;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dC, hC.size() * 4)); CK(cudaMalloc(&dW, hW.size() * 4));
  CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dY0, (size_t)N * O * 4)); CK(cudaMalloc(&dY1, (size_t)N * O * 4));
  CK(cudaMalloc(&dWs, 2ull * K * O * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dC, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  cudaStream_t st = 0;
  CK(cudaFuncSetAttribute(pna::k_linear_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pna::kTsSmem));
  const unsigned grid = (unsigned)((N + pna::kLinM - 1) / pna::kLinM);

  // library kernel (validated): reference result + split weight images in dWs
  if (pna_linear_scaled_fwd(dA, KA, dC, S, dW, dB, dY0, O, N, K, O, dWs, 2ull * K * O * 4, st) != 0) { printf("library call failed: %s\n", pna_last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  pna::k_linear_ts<<<grid, pna::kTsThreads, pna::kTsSmem, st>>>(dA, KA, dC, S, dWs, dB, dY1, O, N, KA);
  CK(cudaDeviceSynchronize());
  std::vector<float> y0((size_t)N * O), y1((size_t)N * O);
  CK(cudaMemcpy(y0.data(), dY0, y0.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(y1.data(), dY1, y1.size() * 4, cudaMemcpyDeviceToHost));
  double maxd = 0, maxref = 0;
  for (size_t i = 0; i < y0.size(); ++i) { maxd = std::max(maxd, (double)std::fabs(y0[i] - y1[i])); maxref = std::max(maxref, (double)std::fabs(y0[i])); }
  // float64 check of a few rows
  double max64 = 0;
  for (long long r = 0; r < N; r += std::max(1ll, N / 64)) {
    for (int o = 0; o < O; ++o) {
      double acc = hB[o];
      for (int s = 0; s < S; ++s)
        for (int k = 0; k < KA; ++k) acc += (double)(hC[r * S + s] * hA[r * KA + k]) * (double)hW[(size_t)o * K + s * KA + k];
      max64 = std::max(max64, std::fabs(acc - (double)y1[r * O + o]));
    }
  }
  printf("N=%lld  max|ts - library| %.3e   max|ts - float64| (sampled rows) %.3e   max|y| %.2f\n", N, maxd, max64, maxref);
  const float t_lib = time_ms(st, 20, [&] { pna_linear_scaled_fwd(dA, KA, dC, S, dW, dB, dY0, O, N, K, O, dWs, 2ull * K * O * 4, st); });
  const float t_ts = time_ms(st, 20, [&] { pna::k_linear_ts<<<grid, pna::kTsThreads, pna::kTsSmem, st>>>(dA, KA, dC, S, dWs, dB, dY1, O, N, KA); });
  printf("library (A in smem, incl. k_split_weight) %.3f ms    TS (A in TMEM) %.3f ms\n", t_lib, t_ts);
  printf(maxd <= 3e-5 * std::max(1.0, maxref) ? "LINEAR_TS OK\n" : "LINEAR_TS MISMATCH\n");
  return 0;
}
