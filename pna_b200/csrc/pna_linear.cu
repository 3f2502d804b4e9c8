// pna_linear_fwd: Y[N, O] = A[N, K] . W[O, K]^T + b in fp32 accuracy on the 5th-generation tensor cores.
//
// This is the first dense linear of the post-aggregation MLP (reference models/pytorch_geometric/pna.py:222-227,
// post_nn[0]; models/dgl/pna_layer.py:31 posttrans), the one place of the PNA layer where tensor cores apply
// (north_star: "the post-MLP uses tensor cores only for its dense linear").  The 1e-5 parity bar rules out plain TF32
// (10-bit mantissa); every operand is therefore split  x = hi + lo  (hi = top 19 bits, lo = x - hi, exact) and three
// tcgen05.mma kind::tf32 products are accumulated in TMEM:  hi.hi + hi.lo + lo.hi  (the dropped lo.lo term is 2^-22).
//
// One CTA per 128-row tile, 10 warps, 3-stage mbarrier ring of {A hi, A lo, W hi, W lo} tiles:
//   warps 0-7  A loaders: coalesced 128-bit loads issued four K blocks ahead, split hi / lo (cvt.rna.tf32: an unbiased
//              split -- truncation accumulates its one-sided error linearly in K) and stored into the 128-byte-swizzled
//              K-major layout UMMA reads (the operand has to pass through registers for the split, so no TMA here).
//              Warps 0-3 are afterwards the epilogue: tcgen05.ld the accumulator, add the bias, store.
//   warp 8     allocates TMEM and issues the MMAs from one elected lane: 12 per 32-wide K block
//              (4 K-steps x 3 products); tcgen05.commit releases the stage / signals the epilogue.
//   warp 9     W producer: the weight is pre-split once per call into the exact swizzled shared-memory image of every
//              K block, so a W tile is one contiguous cp.async.bulk (TMA 1-D) completing on the stage's mbarrier.
//
// Scaled ("compact") mode -- pna_linear_scaled_fwd.  The reference's post-MLP input is cat_s(c_s(i) * agg_i) over the
// degree scalers s (pna.py:247-249): S copies of the same [N, A*F] aggregate, each multiplied by a per-row factor.  In
// this mode A is the COMPACT aggregate (identity scaler only) and the loaders regenerate every scaled copy in registers,
// fl(c_s(i) * a) rounded exactly like the reference's multiply, right before the hi/lo split -- the [N, S*A*F] tensor
// never exists in HBM: the aggregation writes, and this kernel reads, 1/S of the bytes.  K blocks are visited
// compact-block-major, scaler-minor; the W tile of step (kb, s) is block s * (K/32) + kb of the reference's weight.
#include "common.cuh"

namespace pna {

constexpr int kLinM = 128;          // rows per CTA (UMMA_M)
constexpr int kLinBK = 32;          // fp32 per K block = one 128-byte swizzle row
constexpr int lin_stages(int o) { return o <= 128 ? 3 : 2; }   // 64 KB (O=128) / 96 KB (O=256) per stage
constexpr int kLinLoaders = 256;    // warps 0-7: A loaders (warps 0-3 are also the epilogue)
constexpr int kLinThreads = kLinLoaders + 64;   // + warp 8: MMA issuer, warp 9: W tile producer (bulk copies)

__device__ __forceinline__ unsigned lin_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lin_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void lin_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void lin_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lin_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LD_%=;\n\t"
      "bra LW_%=;\n\t"
      "LD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void lin_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
               : "memory");
}
// x -> nearest TF32 value (19 significant bits kept, round to nearest): the split must not be biased, a truncating
// split accumulates its one-sided error linearly in K
__device__ __forceinline__ float lin_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long lin_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3ffffu) >> 4);        // start address, bits [0,14)
  d |= (unsigned long long)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (unsigned long long)(1024 >> 4) << 32;                     // stride byte offset: 8 rows x 128 B
  d |= (unsigned long long)1 << 46;                               // descriptor version (Blackwell)
  d |= (unsigned long long)2 << 61;                               // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void lin_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// byte offset of 16-byte unit j of row r inside a [rows][128 B] tile with the 128-byte swizzle (unit ^= row % 8)
__host__ __device__ __forceinline__ unsigned lin_swz(int r, int j) { return (unsigned)(r * 128 + ((j ^ (r & 7)) << 4)); }

template <int O>   // output width = UMMA_N, a multiple of 16 up to 256
struct LinSmem {
  static constexpr int kSt = lin_stages(O);
  // TMEM accumulators: kMain for hi.hi (K blocks dealt round-robin: fewer truncating accumulation steps and a smaller
  // running sum per accumulator) + 1 for the two cross terms
  static constexpr int kMain = (3 * O <= 512) ? 2 : 1;
  static constexpr unsigned kCols = (kMain + 1) * O <= 64 ? 64 : (kMain + 1) * O <= 128 ? 128 : (kMain + 1) * O <= 256 ? 256 : 512;
  static constexpr int kATile = kLinM * 128;                  // bytes of one A hi (or lo) stage
  static constexpr int kWTile = O * 128;
  static constexpr int kStage = 2 * kATile + 2 * kWTile;
  static constexpr size_t kBytes = 1024 /*align slack*/ + (size_t)kSt * kStage + 256;
};

// Wimg: for every K block the exact shared-memory images of the W hi and W lo tiles ([O rows][128 B], swizzled),
// produced once per call by k_split_weight -- so a tile is ONE contiguous bulk copy (TMA 1-D, no tensor map needed).
template <int O>
__global__ void __launch_bounds__(kLinThreads, 1)
k_linear_3xtf32(const float* __restrict__ A, long long lda, const float* __restrict__ row_scale, int n_rep,
                const float* __restrict__ Wimg, const float* __restrict__ bias, float* __restrict__ Y, long long ldy, long long N,
                int K) {
  constexpr int kSt = LinSmem<O>::kSt;
  extern __shared__ unsigned char lin_raw[];
  const unsigned base = (lin_smem_u32(lin_raw) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
  unsigned char* gbase = lin_raw + (base - lin_smem_u32(lin_raw));
  const unsigned bars = base + kSt * LinSmem<O>::kStage;                    // full[kSt], empty[kSt], tmem_full, tmem slot
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * kLinM;
  const int n_kb = K / kLinBK;          // K blocks of A (compact width)
  const int n_it = n_kb * n_rep;        // pipeline steps = K blocks of W (n_rep == 1 without row scales)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSt; ++s) {
      lin_mbar_init(bars + 8 * s, kLinLoaders / 32 + 1);    // full: one arrive per A-loader warp + the W producer (with tx bytes)
      lin_mbar_init(bars + 8 * (kSt + s), 1);               // empty: tcgen05.commit
    }
    lin_mbar_init(bars + 8 * (2 * kSt), 1);                 // accumulator ready
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {   // TMEM: O columns (power of two >= 32)
    constexpr unsigned cols = LinSmem<O>::kCols;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + 8 * (2 * kSt + 1)), "n"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *reinterpret_cast<volatile unsigned*>(gbase + kSt * LinSmem<O>::kStage + 8 * (2 * kSt + 1));

  if (warp < 8) {
    // ---------------- A loaders: LDG two K blocks ahead -> split hi/lo -> swizzled STS ----------------
    const int tid = threadIdx.x;                            // 0..255
    const int j = tid & 7;                                  // 16-byte unit inside the 128-byte K block row
    const int r_in = tid >> 3;                              // 0..31: row inside a 32-row slab
    constexpr int kSlabs = kLinM / 32;                      // 4
    constexpr int kAhead = 4;                               // K blocks of A kept in flight per thread (16 x 16 B)
    float4 pre[kAhead][kSlabs];
    auto fetch = [&](int kb, float4 (&dst)[kSlabs]) {
#pragma unroll
      for (int sl = 0; sl < kSlabs; ++sl) {
        const long long r = row0 + sl * 32 + r_in;
        dst[sl] = (kb < n_kb && r < N) ? __ldg(reinterpret_cast<const float4*>(A + r * lda + kb * kLinBK + j * 4))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int u = 0; u < kAhead; ++u) fetch(u, pre[u]);
    unsigned phase = 0;
    int s = 0;                                              // ring stage of the next pipeline step
    for (int kb0 = 0; kb0 < n_kb; kb0 += kAhead) {
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {                    // unrolled so that pre[u] stays in registers
        const int kb = kb0 + u;
        if (kb >= n_kb) break;
        for (int rep = 0; rep < n_rep; ++rep) {             // the scaled copies of this K block (one pass without scales)
          float sc[kSlabs];
          if (row_scale) {                                  // issued before the wait: an L1 hit after the first K block
#pragma unroll
            for (int sl = 0; sl < kSlabs; ++sl) {
              const long long r = row0 + sl * 32 + r_in;
              sc[sl] = r < N ? __ldg(row_scale + r * n_rep + rep) : 0.f;
            }
          }
          lin_mbar_wait(bars + 8 * (kSt + s), phase ^ 1);   // stage free
          unsigned char* st = gbase + s * LinSmem<O>::kStage;
#pragma unroll
          for (int sl = 0; sl < kSlabs; ++sl) {
            float4 v = pre[u][sl];
            if (row_scale) {                                // scalers.py: src * scale, rounded to fp32 like the reference
              v.x = __fmul_rn(v.x, sc[sl]); v.y = __fmul_rn(v.y, sc[sl]);
              v.z = __fmul_rn(v.z, sc[sl]); v.w = __fmul_rn(v.w, sc[sl]);
            }
            float4 hi, lo;
            hi.x = lin_tf32(v.x); lo.x = lin_tf32(v.x - hi.x);
            hi.y = lin_tf32(v.y); lo.y = lin_tf32(v.y - hi.y);
            hi.z = lin_tf32(v.z); lo.z = lin_tf32(v.z - hi.z);
            hi.w = lin_tf32(v.w); lo.w = lin_tf32(v.w - hi.w);
            const unsigned off = lin_swz(sl * 32 + r_in, j);  // quarter-warps write whole swizzled 128-byte rows: conflict free
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + LinSmem<O>::kATile + off) = lo;
          }
          if (rep == n_rep - 1) fetch(kb + kAhead, pre[u]);   // refill the slot just consumed
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> visible to the MMA (async proxy)
          __syncwarp();
          if (lane == 0) lin_mbar_arrive(bars + 8 * s);
          if (++s == kSt) { s = 0; phase ^= 1; }
        }
      }
    }
    if (warp < 4) {
      // ---------------- epilogue ----------------
      lin_mbar_wait(bars + 8 * (2 * kSt), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long row = row0 + warp * 32 + lane;        // TMEM lane = accumulator row; warp w owns lanes 32w..32w+31
#pragma unroll
      for (int c0 = 0; c0 < O; c0 += 16) {
        unsigned v[16];
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int acc = 1; acc <= LinSmem<O>::kMain; ++acc) {     // the other main accumulator(s) (only if K reached them) + cross terms
          if (acc < LinSmem<O>::kMain && n_it <= acc) continue;
          unsigned c[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]),
                "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15])
              : "r"(taddr + acc * O));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(c[i]));
        }
        if (row < N) {
          float* yr = Y + row * ldy + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o;
            o.x = __uint_as_float(v[i]) + (bias ? __ldg(bias + c0 + i) : 0.f);
            o.y = __uint_as_float(v[i + 1]) + (bias ? __ldg(bias + c0 + i + 1) : 0.f);
            o.z = __uint_as_float(v[i + 2]) + (bias ? __ldg(bias + c0 + i + 2) : 0.f);
            o.w = __uint_as_float(v[i + 3]) + (bias ? __ldg(bias + c0 + i + 3) : 0.f);
            *reinterpret_cast<float4*>(yr + i) = o;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(O >> 3) << 17) | ((unsigned)(kLinM >> 4) << 24);
    unsigned phase = 0;
    for (int kb = 0; kb < n_it; ++kb) {                     // kb: pipeline step (= K block of the scaled operand)
      const int s = kb % kSt;
      lin_mbar_wait(bars + 8 * s, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const unsigned st = base + s * LinSmem<O>::kStage;
        const unsigned a_hi = st, a_lo = st + LinSmem<O>::kATile, w_hi = st + 2 * LinSmem<O>::kATile,
                       w_lo = w_hi + LinSmem<O>::kWTile;
#pragma unroll
        for (int ks = 0; ks < kLinBK / 8; ++ks) {            // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
          const unsigned ko = ks * 32;
          // the tensor core adds into fp32 with truncation, a bias that grows with the number of accumulation steps
          // times the magnitude of the running sum: the two small cross terms (2^-11 of the main product) get their own
          // accumulator, and the main product alternates between kMain accumulators
          constexpr int kMain = LinSmem<O>::kMain;
          const unsigned corr = tmem + kMain * O, mainacc = tmem + (kb % kMain) * O;
          lin_mma_tf32(corr, lin_desc(a_hi + ko), lin_desc(w_lo + ko), idesc, (kb | ks) ? 1u : 0u);
          lin_mma_tf32(corr, lin_desc(a_lo + ko), lin_desc(w_hi + ko), idesc, 1u);
          lin_mma_tf32(mainacc, lin_desc(a_hi + ko), lin_desc(w_hi + ko), idesc, (kb >= kMain || ks) ? 1u : 0u);
        }
        // release the stage when these MMAs have read it; after the last block also publish the accumulator
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8 * (kSt + s)) : "memory");
        if (kb == n_it - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8 * (2 * kSt)) : "memory");
      }
      __syncwarp();
      if (s == kSt - 1) phase ^= 1;
    }
  } else if (warp == 9 && lane == 0) {
    // ---------------- W tile producer: two bulk copies (hi, lo images) per K block ----------------
    unsigned phase = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kSt;
      const int kb = (it % n_rep) * n_kb + it / n_rep;     // weight K block of (compact block it / n_rep, scaler it % n_rep)
      lin_mbar_wait(bars + 8 * (kSt + s), phase ^ 1);
      const unsigned st = base + s * LinSmem<O>::kStage + 2 * LinSmem<O>::kATile;
      lin_mbar_expect_tx(bars + 8 * s, 2u * LinSmem<O>::kWTile);
      const float* img = Wimg + (long long)kb * (2 * O * kLinBK);
      lin_bulk_g2s(st, img, LinSmem<O>::kWTile, bars + 8 * s);
      lin_bulk_g2s(st + LinSmem<O>::kWTile, img + O * kLinBK, LinSmem<O>::kWTile, bars + 8 * s);
      if (s == kSt - 1) phase ^= 1;
    }
  }
  __syncthreads();
  if (warp == 8) {
    constexpr unsigned cols = LinSmem<O>::kCols;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(cols) : "memory");
  }
}

// W [O, K] -> per K block the swizzled shared-memory images of its hi and lo TF32 parts
__global__ void k_split_weight(const float* __restrict__ W, int O, int K, float* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // one 16-byte unit of W
  const int units_per_row = K / 4;
  if (i >= (long long)O * units_per_row) return;
  const int r = (int)(i / units_per_row), u = (int)(i % units_per_row);
  const int kb = u / 8, j = u % 8;
  const float4 w = *reinterpret_cast<const float4*>(W + (long long)r * K + u * 4);
  float4 hi, lo;
  hi.x = lin_tf32(w.x); lo.x = lin_tf32(w.x - hi.x);
  hi.y = lin_tf32(w.y); lo.y = lin_tf32(w.y - hi.y);
  hi.z = lin_tf32(w.z); lo.z = lin_tf32(w.z - hi.z);
  hi.w = lin_tf32(w.w); lo.w = lin_tf32(w.w - hi.w);
  float* tile = img + (long long)kb * (2 * O * kLinBK);
  const unsigned off = lin_swz(r, j) / 4;
  *reinterpret_cast<float4*>(tile + off) = hi;
  *reinterpret_cast<float4*>(tile + O * kLinBK + off) = lo;
}

template <int O>
static int launch_linear(const float* A, long long lda, const float* row_scale, int n_rep, const float* Wimg, const float* bias,
                         float* Y, long long ldy, long long N, int K, cudaStream_t st) {
  auto kern = k_linear_3xtf32<O>;
  static bool attr_set = false;
  if (!attr_set) {
    PNA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LinSmem<O>::kBytes));
    attr_set = true;
  }
  const long long grid = (N + kLinM - 1) / kLinM;
  kern<<<(unsigned)grid, kLinThreads, LinSmem<O>::kBytes, st>>>(A, lda, row_scale, n_rep, Wimg, bias, Y, ldy, N, K);
  PNA_CUDA_TRY(cudaGetLastError());
  return PNA_OK;
}

}  // namespace pna

using namespace pna;

extern "C" int pna_linear_workspace_bytes(int32_t n_in, int32_t n_out, size_t* bytes) {
  PNA_REQUIRE(bytes != nullptr && n_in > 0 && n_out > 0, PNA_ERR_BAD_ARG, "pna_linear_workspace_bytes: bad argument");
  *bytes = 2ull * (size_t)n_in * (size_t)n_out * sizeof(float);
  return PNA_OK;
}

// a: [n_rows, n_a] (n_a = n_in / n_rep); weight: [n_out, n_in]; row_scale: [n_rows, n_rep] or null (n_rep == 1)
static int linear_common(const float* a, int64_t lda, const float* row_scale, int32_t n_rep, const float* weight, const float* bias,
                         float* y, int64_t ldy, int64_t n_rows, int32_t n_in, int32_t n_out, void* workspace, size_t workspace_bytes,
                         pna_stream_t stream, const char* who) {
  PNA_REQUIRE(n_rows >= 0 && n_in > 0 && n_out > 0 && n_rep >= 1 && n_rep <= PNA_MAX_SCALERS, PNA_ERR_BAD_ARG, "%s: bad sizes", who);
  PNA_REQUIRE(n_in % n_rep == 0 && (n_in / n_rep) % kLinBK == 0, PNA_ERR_UNSUPPORTED,
              "%s: n_in / n_rep must be a multiple of %d", who, kLinBK);
  PNA_REQUIRE(n_out == 64 || n_out == 128 || n_out == 256, PNA_ERR_UNSUPPORTED, "%s: n_out must be 64, 128 or 256", who);
  if (n_rows == 0) return PNA_OK;
  PNA_REQUIRE(a && weight && y && workspace, PNA_ERR_BAD_ARG, "%s: null pointer", who);
  PNA_REQUIRE(workspace_bytes >= 2ull * n_in * n_out * sizeof(float), PNA_ERR_WORKSPACE, "%s: workspace too small", who);
  PNA_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(workspace) |
                reinterpret_cast<uintptr_t>(weight)) & 15u) == 0 && lda % 4 == 0 && ldy % 4 == 0,
              PNA_ERR_UNSUPPORTED, "%s: a, y, weight, workspace must be 16-byte aligned with pitches that are multiples of 4", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* img = static_cast<float*>(workspace);
  const long long units = (long long)n_out * (n_in / 4);
  k_split_weight<<<(unsigned)((units + 255) / 256), 256, 0, st>>>(weight, n_out, n_in, img);
  PNA_CUDA_TRY(cudaGetLastError());
  const int n_a = n_in / n_rep;
  switch (n_out) {
    case 64: return launch_linear<64>(a, lda, row_scale, n_rep, img, bias, y, ldy, n_rows, n_a, st);
    case 128: return launch_linear<128>(a, lda, row_scale, n_rep, img, bias, y, ldy, n_rows, n_a, st);
    default: return launch_linear<256>(a, lda, row_scale, n_rep, img, bias, y, ldy, n_rows, n_a, st);
  }
}

extern "C" int pna_linear_fwd(const float* a, int64_t lda, const float* weight, const float* bias, float* y, int64_t ldy, int64_t n_rows,
                              int32_t n_in, int32_t n_out, void* workspace, size_t workspace_bytes, pna_stream_t stream) {
  return linear_common(a, lda, nullptr, 1, weight, bias, y, ldy, n_rows, n_in, n_out, workspace, workspace_bytes, stream,
                       "pna_linear_fwd");
}

extern "C" int pna_linear_scaled_fwd(const float* a, int64_t lda, const float* row_scale, int32_t n_scalers, const float* weight,
                                     const float* bias, float* y, int64_t ldy, int64_t n_rows, int32_t n_in, int32_t n_out,
                                     void* workspace, size_t workspace_bytes, pna_stream_t stream) {
  PNA_REQUIRE(row_scale != nullptr || n_rows == 0, PNA_ERR_BAD_ARG, "pna_linear_scaled_fwd: row_scale is null");
  return linear_common(a, lda, row_scale, n_scalers, weight, bias, y, ldy, n_rows, n_in, n_out, workspace, workspace_bytes, stream,
                       "pna_linear_scaled_fwd");
}

namespace pna {
__global__ void k_row_scales(const int* __restrict__ rowptr, long long n_rows, int n_scalers, unsigned codes, float avg_log,
                             float avg_lin, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const DegScales ds = deg_scales(__ldg(rowptr + i + 1) - __ldg(rowptr + i), avg_log, avg_lin);
  for (int s = 0; s < n_scalers; ++s) out[i * n_scalers + s] = ds.of((codes >> (4 * s)) & 15u);
}
}  // namespace pna

extern "C" int pna_row_scales(const int32_t* rowptr, int64_t n_rows, int32_t n_scalers, uint32_t scaler_codes, float avg_log,
                              float avg_lin, float* scales, pna_stream_t stream) {
  PNA_REQUIRE(n_rows >= 0 && n_scalers >= 1 && n_scalers <= PNA_MAX_SCALERS, PNA_ERR_BAD_ARG, "pna_row_scales: bad sizes");
  for (int s = 0; s < n_scalers; ++s)
    PNA_REQUIRE(((scaler_codes >> (4 * s)) & 15u) <= PNA_SCALE_INVERSE_LINEAR, PNA_ERR_BAD_ARG, "pna_row_scales: bad scaler code");
  if (n_rows == 0) return PNA_OK;
  PNA_REQUIRE(rowptr && scales, PNA_ERR_BAD_ARG, "pna_row_scales: null pointer");
  k_row_scales<<<(unsigned)((n_rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, n_rows, n_scalers, scaler_codes,
                                                                                            avg_log, avg_lin, scales);
  PNA_CUDA_TRY(cudaGetLastError());
  return PNA_OK;
}
