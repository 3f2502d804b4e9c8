// Test infrastructure: run CUDA kernels that have NO intra-block communication (no shared memory, no shuffles, no
// barriers) on the host, one thread after the other, so that their index arithmetic and control flow can be checked in
// the CPU suite.  tests/emu/build_emu.py rewrites `kernel<<<grid, block, smem, stream>>>(args)` into EMU_LAUNCH(...) and
// compiles the translation unit with g++ against this header.  Not part of the product; never loaded by pna_b200/.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
}
static thread_local emu::Idx threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

template <class F>
static void emu_launch(F body, dim3 grid, dim3 block) {
  gridDim = grid; blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
              body();
            }
      }
}
#define EMU_LAUNCH(kern, grid, block, ...) emu_launch([&]() { kern(__VA_ARGS__); }, dim3(grid), dim3(block))

#define cudaGetLastError() cudaSuccess
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#define cudaDeviceGetAttribute(p, attr, dev) (*(p) = 2, cudaSuccess)     /* "2 SMs": small grids, grid-stride loops get exercised */
#include <stdlib.h>
static inline void emu_unsupported() { abort(); }                        /* inline PTX: that kernel is not emulated */
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline void __stwt(T* p, T v) { *p = v; }
// compiled with -ffp-contract=off: one rounding per operation, like the _rn intrinsics
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
static inline float4 atomicAdd(float4* p, float4 v) {
  const float4 o = *p;
  p->x = o.x + v.x; p->y = o.y + v.y; p->z = o.z + v.z; p->w = o.w + v.w;
  return o;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
