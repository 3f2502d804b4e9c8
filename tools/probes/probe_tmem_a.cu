// Probe for round 2: tcgen05.mma with the A operand in TENSOR MEMORY (TS form), kind::tf32.
//
// Question it answers on a B200 before the post-linear kernel is rewritten around it:
//   * does  tcgen05.mma.cta_group::1.kind::tf32 [d], [a_tmem], b_desc, idesc, p  assemble and run for sm_100a,
//   * which TMEM cell holds A[m][k]: hypothesis  lane = m, column = a_col0 + k  (one 32-bit cell per tf32 element),
//   * how a 32-wide K block is walked: hypothesis  K step ks (8 elements) reads columns a_col0 + 8*ks ...
//   * cycles per MMA for the SS form (A in shared memory, what pna_linear.cu does today) and the TS form.
// B is the identity on the first 32 of its 128 rows (K-major, SWIZZLE_128B -- the validated layout of pna_linear.cu), so
// D[m][n] = A[m][n] for n < 32 and 0 for n >= 32.
//
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe tools/probes/probe_tmem_a.cu && /tmp/probe
// Not part of libpna_sm100.so; not run by the tests.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned swz(int r, int j) { return (unsigned)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ unsigned long long desc_sw128(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3ffffu) >> 4);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int M = 128, N = 128, KB = 32;
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);

// mode 0: A[m][k] = m   mode 1: A[m][k] = k      ts != 0: A from TMEM, else A from shared memory (reference behaviour)
// reps > 1: timing loop (the same K block again and again, accumulate on)
__global__ void __launch_bounds__(160, 1) k_probe(float* __restrict__ out, long long* __restrict__ cycles, int mode, int ts, int reps) {
  extern __shared__ unsigned char raw[];
  const unsigned base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char* g = raw + (base - smem_u32(raw));
  unsigned char* sA = g;                 // [128][128 B] swizzled
  unsigned char* sB = g + M * 128;       // [128][128 B] swizzled
  const unsigned bar = base + 2 * M * 128;          // mbarrier
  const unsigned slot = bar + 8;                    // TMEM base address
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // operands in shared memory (B always; A for the SS form)
  for (int i = threadIdx.x; i < M * 8; i += blockDim.x) {
    const int r = i >> 3, j = i & 7;
    float4 a, b;
    const int k0 = j * 4;
    a.x = mode ? (float)(k0 + 0) : (float)r; a.y = mode ? (float)(k0 + 1) : (float)r;
    a.z = mode ? (float)(k0 + 2) : (float)r; a.w = mode ? (float)(k0 + 3) : (float)r;
    b.x = (r == k0 + 0) ? 1.f : 0.f; b.y = (r == k0 + 1) ? 1.f : 0.f; b.z = (r == k0 + 2) ? 1.f : 0.f; b.w = (r == k0 + 3) ? 1.f : 0.f;
    *reinterpret_cast<float4*>(sA + swz(r, j)) = a;
    *reinterpret_cast<float4*>(sB + swz(r, j)) = b;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *reinterpret_cast<volatile unsigned*>(g + 2 * M * 128 + 8);
  const unsigned d_col = 0, a_col = 128;           // D: columns [0,128)   A: columns [128,160)

  if (warp < 4 && ts) {
    // thread (warp w, lane l) owns TMEM lane 32w + l = row m; 32 consecutive columns = the 32 K elements of the block
    const int m = warp * 32 + lane;
    unsigned v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(mode ? (float)k : (float)m);
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + a_col;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 4 && lane == 0) {
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < KB / 8; ++ks) {
        const unsigned acc = (rep | ks) ? 1u : 0u;
        const unsigned long long bdesc = desc_sw128(smem_u32(sB) + ks * 32);
        if (ts) {
          const unsigned a_t = tmem + a_col + ks * 8;          // hypothesis: 8 columns per K step
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + d_col), "r"(a_t), "l"(bdesc), "r"(kIdesc), "r"(acc)
              : "memory");
        } else {
          const unsigned long long adesc = desc_sw128(smem_u32(sA) + ks * 32);
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + d_col), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(acc)
              : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int m = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
      unsigned v[16];
      const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + d_col + (unsigned)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) out[m * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

int main() {
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, M * N * sizeof(float)));
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  const size_t smem = 1024 + 2 * M * 128 + 64;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> h(M * N);
  bool all_ok = true;
  for (int ts = 0; ts < 2; ++ts)
    for (int mode = 0; mode < 2; ++mode) {
      CK(cudaMemset(d_out, 0xff, M * N * sizeof(float)));
      k_probe<<<1, 160, smem>>>(d_out, d_cyc, mode, ts, 1);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_out, M * N * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0, first_m = -1, first_n = -1;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          const float want = n < KB ? (mode ? (float)n : (float)m) : 0.f;
          if (h[m * N + n] != want) { if (!bad) { first_m = m; first_n = n; } ++bad; }
        }
      printf("%s form, A[m][k] = %s: %s", ts ? "TS (A in TMEM)" : "SS (A in smem)", mode ? "k" : "m", bad ? "MISMATCH" : "ok");
      if (bad) printf("  (%d cells; first D[%d][%d] = %g)", bad, first_m, first_n, h[first_m * N + first_n]);
      printf("\n");
      all_ok &= !bad;
      if (bad) {   // dump the corner so that the actual layout can be read off
        for (int m = 0; m < 4; ++m) { for (int n = 0; n < 12; ++n) printf("%6g ", h[m * N + n]); printf("  | row %d\n", m); }
        for (int m = 32; m < 34; ++m) { for (int n = 0; n < 12; ++n) printf("%6g ", h[m * N + n]); printf("  | row %d\n", m); }
      }
    }
  for (int ts = 0; ts < 2; ++ts) {
    long long cyc = 0;
    const int reps = 256;
    k_probe<<<1, 160, smem>>>(d_out, d_cyc, 0, ts, reps);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    printf("%s: %lld cycles for %d MMAs (128x128x8 tf32) = %.1f cycles / MMA\n", ts ? "TS" : "SS", cyc, reps * 4, (double)cyc / (reps * 4));
  }
  printf(all_ok ? "PROBE OK\n" : "PROBE: layout hypothesis wrong somewhere, see dumps\n");
  return 0;
}
