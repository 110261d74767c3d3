// Shared helpers for the progan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/progan_b200.h"

namespace pg {

void set_error(const char *fmt, ...);

#define PG_CHECK_ARG(cond, ...)                                                  \
  do {                                                                           \
    if (!(cond)) {                                                               \
      pg::set_error(__VA_ARGS__);                                                \
      return PG_ERR_INVALID;                                                     \
    }                                                                            \
  } while (0)

#define PG_CHECK_LAUNCH(name)                                                    \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      pg::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
      return PG_ERR_CUDA;                                                        \
    }                                                                            \
    return PG_OK;                                                                \
  } while (0)

// dispatch on the activation dtype code
#define PG_DISPATCH_DTYPE(dtype, T, ...)                                         \
  do {                                                                           \
    if ((dtype) == PG_F32) {                                                     \
      using T = float;                                                           \
      __VA_ARGS__;                                                               \
    } else if ((dtype) == PG_BF16) {                                             \
      using T = __nv_bfloat16;                                                   \
      __VA_ARGS__;                                                               \
    } else {                                                                     \
      pg::set_error("unknown dtype code %d", (int)(dtype));                      \
      return PG_ERR_INVALID;                                                     \
    }                                                                            \
  } while (0)

__device__ __forceinline__ float ldf(const float *p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// 8-element vector access (bf16: one 16-byte transaction, f32: two).
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const float *p) {
  F8 o;
  float4 a = *reinterpret_cast<const float4 *>(p);
  float4 b = *reinterpret_cast<const float4 *>(p + 4);
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  return o;
}
__device__ __forceinline__ F8 ld8(const __nv_bfloat16 *p) {
  F8 o;
  uint4 raw = *reinterpret_cast<const uint4 *>(p);
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    o.v[2 * i] = f.x;
    o.v[2 * i + 1] = f.y;
  }
  return o;
}
__device__ __forceinline__ void st8(float *p, const F8 &o) {
  *reinterpret_cast<float4 *>(p) = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
  *reinterpret_cast<float4 *>(p + 4) = make_float4(o.v[4], o.v[5], o.v[6], o.v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16 *p, const F8 &o) {
  uint4 raw;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o.v[2 * i], o.v[2 * i + 1]);
  *reinterpret_cast<uint4 *>(p) = raw;
}

// raw (unconverted) 8-element vectors: issue the loads first, convert later (more bytes in flight)
struct Raw8f { float4 a, b; };
__device__ __forceinline__ Raw8f ldraw8(const float *p) {
  Raw8f r;
  r.a = *reinterpret_cast<const float4 *>(p);
  r.b = *reinterpret_cast<const float4 *>(p + 4);
  return r;
}
__device__ __forceinline__ uint4 ldraw8(const __nv_bfloat16 *p) {
  return *reinterpret_cast<const uint4 *>(p);
}
__device__ __forceinline__ F8 unpack8(const Raw8f &r) {
  F8 o;
  o.v[0] = r.a.x; o.v[1] = r.a.y; o.v[2] = r.a.z; o.v[3] = r.a.w;
  o.v[4] = r.b.x; o.v[5] = r.b.y; o.v[6] = r.b.z; o.v[7] = r.b.w;
  return o;
}
__device__ __forceinline__ F8 unpack8(const uint4 &raw) {
  F8 o;
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    o.v[2 * i] = f.x;
    o.v[2 * i + 1] = f.y;
  }
  return o;
}
template <typename T> struct RawOf;
template <> struct RawOf<float> { using type = Raw8f; };
template <> struct RawOf<__nv_bfloat16> { using type = uint4; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread gets the result. `red` needs 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization
// and starts with grid_dep_sync(): griddepcontrol.wait blocks until the grids it depends on have
// completed and flushed (a no-op for a launch without the attribute), launch_dependents then lets
// the NEXT kernel of the stream be scheduled while this one runs - its CTAs take the SM slots this
// grid frees and sit at their own wait - so the ~2 us launch/drain gap between two dependent
// kernels (850 us of idle SMs per 128 px iteration, 370 kernels) overlaps the tail of the
// predecessor.  Captured into a CUDA graph the attribute becomes a programmatic dependency edge.
// Only SMALL launches get the attribute: the pre-launched CTAs of a machine-filling grid occupy
// SM slots while they wait, which takes them away from the kernels of the concurrent streams
// (weight-gradient side stream, gradient-penalty chain) - measured +1.5 % step time at 128 px with
// the attribute on every launch, -4..5 % at 8-32 px where every kernel is small.
// PG_PDL=0 turns the attribute off, PG_PDL=2 puts it on every launch.
__device__ __forceinline__ void grid_dep_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

int pdl_mode();          // 0 off, 1 small launches only (default), 2 every launch
static inline int sm_count();
static inline bool pdl_for(long long ctas) {
  const int m = pdl_mode();
  return m == 2 || (m == 1 && ctas <= 2ll * sm_count());
}

template <typename F>
struct Launcher {
  F kern;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  template <typename... A>
  cudaError_t operator()(A &&...a) {
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<A &&>(a)...);
  }
};
template <typename F>
static inline Launcher<F> launcher(F kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   long long work_ctas = -1) {
  // work_ctas: for persistent kernels (grid = #SMs) the number of CTA-sized work items instead
  Launcher<F> l;
  l.kern = kern;
  l.cfg = cudaLaunchConfig_t{};
  l.cfg.gridDim = grid;
  l.cfg.blockDim = block;
  l.cfg.dynamicSmemBytes = smem;
  l.cfg.stream = stream;
  l.at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  const long long ctas = work_ctas >= 0 ? work_ctas : (long long)grid.x * grid.y * grid.z;
  l.at[0].val.programmaticStreamSerializationAllowed = pdl_for(ctas) ? 1 : 0;
  return l;
}

static inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// grid for a grid-stride bandwidth kernel: a multiple of the SM count.
static inline int bw_grid(long long work_items, int per_block, int ctas_per_sm = 8) {
  long long need = (work_items + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace pg
