// Resampling, fade-in blend and small elementwise kernels.
//
// Reference: F.interpolate(scale_factor=2 / 0.5, mode='bilinear', align_corners=False)
// at progan_modules.py:168,205,299,303 and the alpha blend at :212,305.
//   x0.5  == 2x2 average pool (exactly);
//   x2    == separable 2-tap stencil  o[2i]   = .25 x[max(i-1,0)] + .75 x[i]
//                                     o[2i+1] = .75 x[i] + .25 x[min(i+1,n-1)]
// (SURVEY.md F1, Appendix D).  All are HBM-bound: 16-byte vector accesses over the
// channel dimension, grid sized in multiples of the SM count.
#include "common.cuh"
#include <type_traits>

namespace pg {

// G = channel group width handled by one thread (8 = vector path, 1 = scalar path)
template <typename T, int G>
struct Grp {
  float v[G];
  __device__ __forceinline__ void load(const T *p) {
    if constexpr (G == 8) {
      F8 t = ld8(p);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = t.v[e];
    } else {
      v[0] = ldf(p);
    }
  }
  __device__ __forceinline__ void store(T *p) const {
    if constexpr (G == 8) {
      F8 t;
#pragma unroll
      for (int e = 0; e < 8; ++e) t.v[e] = v[e];
      st8(p, t);
    } else {
      stf(p, v[0]);
    }
  }
};

template <typename T, int G, typename I>
__global__ void __launch_bounds__(256)
avgpool2_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int C) {
  pg::grid_dep_sync();
  const int Ho = H / 2, Wo = W / 2, cg = C / G;
  const I total = (I)N * (I)Ho * (I)Wo * (I)cg;
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total;
       i += (I)gridDim.x * (I)blockDim.x) {
    I t = i / (I)cg;
    const int c = (int)(i - t * (I)cg) * G;
    I t2 = t / (I)Wo;
    const int ox = (int)(t - t2 * (I)Wo);
    const long long n = (long long)(t2 / (I)Ho);
    const int oy = (int)(t2 - (I)n * (I)Ho);
    const T *base = x + ((n * H + 2 * oy) * W + 2 * ox) * (long long)C + c;
    Grp<T, G> a, b, cc, d, o;
    a.load(base);
    b.load(base + C);
    cc.load(base + (long long)W * C);
    d.load(base + (long long)W * C + C);
#pragma unroll
    for (int e = 0; e < G; ++e) o.v[e] = 0.25f * (a.v[e] + b.v[e] + cc.v[e] + d.v[e]);
    o.store(y + ((n * Ho + oy) * Wo + ox) * (long long)C + c);
  }
}

template <typename T, int G, typename I>
__global__ void __launch_bounds__(256)
avgpool2_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int N, int H, int W, int C) {
  pg::grid_dep_sync();
  const int Ho = H / 2, Wo = W / 2, cg = C / G;
  const I total = (I)N * (I)H * (I)W * (I)cg;
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total;
       i += (I)gridDim.x * (I)blockDim.x) {
    I t = i / (I)cg;
    const int c = (int)(i - t * (I)cg) * G;
    I t2 = t / (I)W;
    const int xx = (int)(t - t2 * (I)W);
    const long long n = (long long)(t2 / (I)H);
    const int yy = (int)(t2 - (I)n * (I)H);
    Grp<T, G> g;
    g.load(dy + ((n * Ho + yy / 2) * Wo + xx / 2) * (long long)C + c);
#pragma unroll
    for (int e = 0; e < G; ++e) g.v[e] *= 0.25f;
    g.store(dx + (long long)i * G);
  }
}

// C == 1 planes (an NCHW fp32 image passed as N*K planes): four input columns per thread with
// 16-byte accesses instead of one scalar per thread.
__global__ void __launch_bounds__(256)
avgpool2_kernel_plane(const float *__restrict__ x, float *__restrict__ y, long long planes, int H,
                      int W) {
  pg::grid_dep_sync();
  const int Ho = H / 2, Wq = W / 4;
  const long long total = planes * Ho * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Wq);
    const long long t = i / Wq;
    const int oy = (int)(t % Ho);
    const long long pl = t / Ho;
    const float *r0 = x + (pl * H + 2 * oy) * W + 4 * q;
    const float4 a = *reinterpret_cast<const float4 *>(r0);
    const float4 b = *reinterpret_cast<const float4 *>(r0 + W);
    float2 o;
    o.x = 0.25f * (a.x + a.y + b.x + b.y);
    o.y = 0.25f * (a.z + a.w + b.z + b.w);
    *reinterpret_cast<float2 *>(y + (pl * Ho + oy) * (W / 2) + 2 * q) = o;
  }
}

__global__ void __launch_bounds__(256)
avgpool2_bwd_kernel_plane(const float *__restrict__ dy, float *__restrict__ dx, long long planes,
                          int H, int W) {
  pg::grid_dep_sync();
  const int Wq = W / 4;
  const long long total = planes * H * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Wq);
    const long long t = i / Wq;
    const int yy = (int)(t % H);
    const long long pl = t / H;
    const float2 g = *reinterpret_cast<const float2 *>(dy + (pl * (H / 2) + yy / 2) * (W / 2) + 2 * q);
    *reinterpret_cast<float4 *>(dx + (pl * H + yy) * W + 4 * q) =
        make_float4(0.25f * g.x, 0.25f * g.x, 0.25f * g.y, 0.25f * g.y);
  }
}

// Work decomposition of the two x2 kernels: a block owns a compact 2-D tile of pixels (TW wide,
// 256 / (cg * TW) high, all channel groups), so the rows its neighbours share are re-read from L1
// instead of L2 (in grid-stride LINEAR order every input row was fetched from L2 by several
// blocks: 4x the input bytes on the L2 -> SM path, 0.34-0.45 of the HBM roofline).
struct Tile2D {
  int n, y, x, c;
  bool ok;
};
template <int G>
__device__ __forceinline__ Tile2D tile_item(long long tile, int tid, int Hd, int Wd, int C, int TW,
                                            int TH) {
  const int cg = C / G;
  const int tiles_x = (Wd + TW - 1) / TW, tiles_y = (Hd + TH - 1) / TH;
  const int tx = (int)(tile % tiles_x);
  const long long t2 = tile / tiles_x;
  const int ty = (int)(t2 % tiles_y);
  Tile2D r;
  r.n = (int)(t2 / tiles_y);
  const int c = tid % cg, pp = tid / cg;
  r.c = c * G;
  r.x = tx * TW + pp % TW;
  r.y = ty * TH + pp / TW;
  r.ok = pp < TW * TH && r.x < Wd && r.y < Hd;
  return r;
}

// Output quad (rows 2i+1, 2i+2) x (cols 2j+1, 2j+2) depends on the 2x2 input (i..i+1, j..j+1)
// only (indices clamped; i, j run from -1): four loads, up to four stores per item.
template <typename T, int G>
__global__ void __launch_bounds__(256)
upsample2_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int C, int TW,
                 int TH, long long tiles) {
  pg::grid_dep_sync();
  const int Ho = 2 * H, Wo = 2 * W;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const Tile2D it = tile_item<G>(tile, threadIdx.x, H + 1, W + 1, C, TW, TH);
    if (!it.ok) continue;
    const int i = it.y - 1, j = it.x - 1;
    const int y0 = max(i, 0), y1 = min(i + 1, H - 1), x0 = max(j, 0), x1 = min(j + 1, W - 1);
    const T *b = x + (long long)it.n * H * W * C + it.c;
    Grp<T, G> a00, a01, a10, a11;
    a00.load(b + ((long long)y0 * W + x0) * C);
    a01.load(b + ((long long)y0 * W + x1) * C);
    a10.load(b + ((long long)y1 * W + x0) * C);
    a11.load(b + ((long long)y1 * W + x1) * C);
    T *o = y + (long long)it.n * Ho * Wo * C + it.c;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int oy = 2 * i + 1 + dy;
      if (oy < 0 || oy >= Ho) continue;
      const float wy0 = dy ? 0.25f : 0.75f, wy1 = 1.f - wy0;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ox = 2 * j + 1 + dx;
        if (ox < 0 || ox >= Wo) continue;
        const float wx0 = dx ? 0.25f : 0.75f, wx1 = 1.f - wx0;
        Grp<T, G> r;
#pragma unroll
        for (int e = 0; e < G; ++e)
          r.v[e] = wy0 * (wx0 * a00.v[e] + wx1 * a01.v[e]) + wy1 * (wx0 * a10.v[e] + wx1 * a11.v[e]);
        r.store(o + ((long long)oy * Wo + ox) * C);
      }
    }
  }
}

// transpose of the stencil: dx[i] = sum over o in {2i-1,2i,2i+1,2i+2} clamped to [0,2n-1]
// with weights {.25,.75,.75,.25} per axis.
template <typename T, int G>
__global__ void __launch_bounds__(256)
upsample2_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int N, int H, int W, int C,
                     int TW, int TH, long long tiles) {
  pg::grid_dep_sync();
  const int Ho = 2 * H, Wo = 2 * W;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const Tile2D it = tile_item<G>(tile, threadIdx.x, H, W, C, TW, TH);
    if (!it.ok) continue;
    const T *b = dy + (long long)it.n * Ho * Wo * C + it.c;
    // all sixteen taps are loaded (raw) before the first multiply: the kernel lives on L1/L2 hits
    // and needs the loads in flight, not the arithmetic
    typename std::conditional<G == 8, typename RawOf<T>::type, float>::type raw[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = min(max(2 * it.y - 1 + a, 0), Ho - 1);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int ox = min(max(2 * it.x - 1 + bb, 0), Wo - 1);
        const T *src = b + ((long long)oy * Wo + ox) * C;
        if constexpr (G == 8) raw[a][bb] = ldraw8(src);
        else raw[a][bb] = ldf(src);
      }
    }
    float acc[G];
#pragma unroll
    for (int e = 0; e < G; ++e) acc[e] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float wgt = wt[a] * wt[bb];
        if constexpr (G == 8) {
          const F8 v = unpack8(raw[a][bb]);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, v.v[e], acc[e]);
        } else {
          acc[0] = fmaf(wgt, raw[a][bb], acc[0]);
        }
      }
    Grp<T, G> o;
#pragma unroll
    for (int e = 0; e < G; ++e) o.v[e] = acc[e];
    o.store(dx + (((long long)it.n * H + it.y) * W + it.x) * C + it.c);
  }
}

// out = ca*a + cb*b with ca = a0 + a1*alpha, cb = b0 + b1*alpha (b may be null)
template <typename T>
__global__ void __launch_bounds__(256)
axpby_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, long long n,
             float a0, float a1, float b0, float b1, const float *__restrict__ alpha_dev) {
  pg::grid_dep_sync();
  const float alpha = alpha_dev ? *alpha_dev : 0.f;
  const float ca = a0 + a1 * alpha, cb = b0 + b1 * alpha;
  const long long nv = n >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    F8 va = ld8(a + i * 8), o;
    if (b) {
      F8 vb = ld8(b + i * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] = ca * va.v[e] + cb * vb.v[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] = ca * va.v[e];
    }
    st8(out + i * 8, o);
  }
  // scalar tail
  for (long long i = (nv << 3) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = ca * ldf(a + i);
    if (b) v += cb * ldf(b + i);
    stf(out + i, v);
  }
}

__global__ void __launch_bounds__(256)
tanh_kernel(const float *__restrict__ x, const float *__restrict__ y_saved,
            float *__restrict__ out, long long n, int bwd) {
  pg::grid_dep_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (!bwd) out[i] = tanhf(x[i]);
    else { const float y = y_saved[i]; out[i] = x[i] * (1.f - y * y); }
  }
}

}  // namespace pg

using namespace pg;

#define PG_RESAMPLE_ENTRY(fn, kernel, work_expr, even_check)                                    \
  extern "C" int fn(const void *x, void *y, int N, int H, int W, int C, int dtype,             \
                    void *stream) {                                                            \
    PG_CHECK_ARG(x && y, #fn ": null pointer");                                                \
    PG_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, #fn ": bad dims");                          \
    PG_CHECK_ARG(!(even_check) || (H % 2 == 0 && W % 2 == 0), #fn ": H and W must be even");   \
    const long long work = (work_expr);                                                        \
    if (C == 1 && dtype == PG_F32 && W % 4 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) { \
      const long long items = (long long)N * H * W / 4 / ((even_check) == 1 ? 2 : 1);          \
      pg::launcher(kernel##_plane, bw_grid(items, 256), 256, 0, (cudaStream_t)stream)(                    \
          (const float *)x, (float *)y, (long long)N, H, W);                                   \
      PG_CHECK_LAUNCH(#fn);                                                                    \
    }                                                                                          \
    /* 32-bit index arithmetic whenever it fits: the 64-bit div/mod chain per element made   \
       these kernels ALU-bound */                                                              \
    const bool small = work * 4 < (1ll << 31);                                                 \
    if (C % 8 == 0) {                                                                          \
      const int grid = bw_grid(work / 8, 256);                                                 \
      PG_DISPATCH_DTYPE(dtype, T, {                                                            \
        if (small) pg::launcher(kernel<T, 8, unsigned>, grid, 256, 0, (cudaStream_t)stream)(             \
                       (const T *)x, (T *)y, N, H, W, C);                                      \
        else pg::launcher(kernel<T, 8, long long>, grid, 256, 0, (cudaStream_t)stream)(                  \
                 (const T *)x, (T *)y, N, H, W, C);                                            \
      });                                                                                      \
    } else {                                                                                   \
      const int grid = bw_grid(work, 256);                                                     \
      PG_DISPATCH_DTYPE(dtype, T, {                                                            \
        if (small) pg::launcher(kernel<T, 1, unsigned>, grid, 256, 0, (cudaStream_t)stream)(             \
                       (const T *)x, (T *)y, N, H, W, C);                                      \
        else pg::launcher(kernel<T, 1, long long>, grid, 256, 0, (cudaStream_t)stream)(                  \
                 (const T *)x, (T *)y, N, H, W, C);                                            \
      });                                                                                      \
    }                                                                                          \
    PG_CHECK_LAUNCH(#fn);                                                                      \
  }

PG_RESAMPLE_ENTRY(pg_avgpool2, avgpool2_kernel, (long long)N * (H / 2) * (W / 2) * C, 1)
PG_RESAMPLE_ENTRY(pg_avgpool2_bwd, avgpool2_bwd_kernel, (long long)N * H * W * C, 2)
// tile shape for the x2 kernels: `Hd x Wd` = pixel domain the items live on
template <int G>
static void tile_shape(int Hd, int Wd, int C, int *TW, int *TH) {
  const int cg = C / G;
  int pix = 256 / cg;
  if (pix < 1) pix = 1;
  int tw = 8;
  while (tw > Wd || tw > pix) tw >>= 1;
  if (tw < 1) tw = 1;
  *TW = tw;
  *TH = pix / tw > 0 ? pix / tw : 1;
}

#define PG_X2_ENTRY(fn, kernel, HD, WD)                                                          \
  extern "C" int fn(const void *x, void *y, int N, int H, int W, int C, int dtype,             \
                    void *stream) {                                                            \
    PG_CHECK_ARG(x && y, #fn ": null pointer");                                                \
    PG_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, #fn ": bad dims");                          \
    int TW, TH;                                                                                \
    if (C % 8 == 0 && C / 8 <= 256) {                                                          \
      tile_shape<8>(HD, WD, C, &TW, &TH);                                                      \
      const long long tiles = (long long)N * (((HD) + TH - 1) / TH) * (((WD) + TW - 1) / TW);  \
      const int grid = bw_grid(tiles, 1);                                                      \
      PG_DISPATCH_DTYPE(dtype, T, (pg::launcher(kernel<T, 8>, grid, 256, 0, (cudaStream_t)stream)(       \
                                      (const T *)x, (T *)y, N, H, W, C, TW, TH, tiles)));      \
    } else {                                                                                   \
      PG_CHECK_ARG(C <= 256, #fn ": C must be a multiple of 8 or <= 256");                     \
      tile_shape<1>(HD, WD, C, &TW, &TH);                                                      \
      const long long tiles = (long long)N * (((HD) + TH - 1) / TH) * (((WD) + TW - 1) / TW);  \
      const int grid = bw_grid(tiles, 1);                                                      \
      PG_DISPATCH_DTYPE(dtype, T, (pg::launcher(kernel<T, 1>, grid, 256, 0, (cudaStream_t)stream)(       \
                                      (const T *)x, (T *)y, N, H, W, C, TW, TH, tiles)));      \
    }                                                                                          \
    PG_CHECK_LAUNCH(#fn);                                                                      \
  }

PG_X2_ENTRY(pg_upsample2, upsample2_kernel, H + 1, W + 1)
PG_X2_ENTRY(pg_upsample2_bwd, upsample2_bwd_kernel, H, W)

extern "C" int pg_blend(const void *a, const void *b, void *out, long long n,
                        const float *alpha_dev, int dtype, void *stream) {
  PG_CHECK_ARG(a && b && out && alpha_dev, "pg_blend: null pointer");
  PG_CHECK_ARG(n > 0, "pg_blend: n must be > 0");
  const int grid = bw_grid((n + 7) / 8, 256);
  PG_DISPATCH_DTYPE(dtype, T, pg::launcher(axpby_kernel<T>, grid, 256, 0, (cudaStream_t)stream)(
                                  (const T *)a, (const T *)b, (T *)out, n, 1.f, -1.f, 0.f, 1.f,
                                  alpha_dev));
  PG_CHECK_LAUNCH("pg_blend");
}

extern "C" int pg_scale(const void *x, void *out, long long n, float c0, float c1,
                        const float *alpha_dev, int dtype, void *stream) {
  PG_CHECK_ARG(x && out, "pg_scale: null pointer");
  PG_CHECK_ARG(n > 0, "pg_scale: n must be > 0");
  PG_CHECK_ARG(alpha_dev || c1 == 0.f, "pg_scale: c1 != 0 needs alpha_dev");
  const int grid = bw_grid((n + 7) / 8, 256);
  PG_DISPATCH_DTYPE(dtype, T, pg::launcher(axpby_kernel<T>, grid, 256, 0, (cudaStream_t)stream)(
                                  (const T *)x, (const T *)nullptr, (T *)out, n, c0, c1, 0.f,
                                  0.f, alpha_dev));
  PG_CHECK_LAUNCH("pg_scale");
}

extern "C" int pg_tanh_fwd(const float *x, float *y, long long n, void *stream) {
  PG_CHECK_ARG(x && y && n > 0, "pg_tanh_fwd: bad args");
  pg::launcher(tanh_kernel, bw_grid(n, 256), 256, 0, (cudaStream_t)stream)(x, nullptr, y, n, 0);
  PG_CHECK_LAUNCH("pg_tanh_fwd");
}

extern "C" int pg_tanh_bwd(const float *dy, const float *y, float *dx, long long n,
                           void *stream) {
  PG_CHECK_ARG(dy && y && dx && n > 0, "pg_tanh_bwd: bad args");
  pg::launcher(tanh_kernel, bw_grid(n, 256), 256, 0, (cudaStream_t)stream)(dy, y, dx, n, 1);
  PG_CHECK_LAUNCH("pg_tanh_bwd");
}
