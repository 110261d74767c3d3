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

template <typename T, int G, typename I>
__global__ void __launch_bounds__(256)
upsample2_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, cg = C / G;
  const I total = (I)N * (I)Ho * (I)Wo * (I)cg;
  for (I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i < total;
       i += (I)gridDim.x * (I)blockDim.x) {
    I t = i / (I)cg;
    const int c = (int)(i - t * (I)cg) * G;
    I t2 = t / (I)Wo;
    const int ox = (int)(t - t2 * (I)Wo);
    const long long n = (long long)(t2 / (I)Ho);
    const int oy = (int)(t2 - (I)n * (I)Ho);
    int y0, y1, x0, x1;
    float wy0, wx0;
    if (oy & 1) { y0 = oy >> 1; y1 = min(y0 + 1, H - 1); wy0 = 0.75f; }
    else        { y1 = oy >> 1; y0 = max(y1 - 1, 0);     wy0 = 0.25f; }
    if (ox & 1) { x0 = ox >> 1; x1 = min(x0 + 1, W - 1); wx0 = 0.75f; }
    else        { x1 = ox >> 1; x0 = max(x1 - 1, 0);     wx0 = 0.25f; }
    const float wy1 = 1.f - wy0, wx1 = 1.f - wx0;
    const T *b = x + n * H * W * (long long)C + c;
    Grp<T, G> a00, a01, a10, a11, o;
    a00.load(b + ((long long)y0 * W + x0) * C);
    a01.load(b + ((long long)y0 * W + x1) * C);
    a10.load(b + ((long long)y1 * W + x0) * C);
    a11.load(b + ((long long)y1 * W + x1) * C);
#pragma unroll
    for (int e = 0; e < G; ++e)
      o.v[e] = wy0 * (wx0 * a00.v[e] + wx1 * a01.v[e]) + wy1 * (wx0 * a10.v[e] + wx1 * a11.v[e]);
    o.store(y + (long long)i * G);
  }
}

// transpose of the stencil: dx[i] = sum over o in {2i-1,2i,2i+1,2i+2} clamped to [0,2n-1]
// with weights {.25,.75,.75,.25} per axis.
template <typename T, int G, typename I>
__global__ void __launch_bounds__(256)
upsample2_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int N, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, cg = C / G;
  const long long total = (long long)N * H * W * cg;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * G;
    long long t = i / cg;
    const int xx = (int)(t % W);
    t /= W;
    const int yy = (int)(t % H);
    const long long n = t / H;
    const T *b = dy + n * Ho * Wo * (long long)C + c;
    float acc[G];
#pragma unroll
    for (int e = 0; e < G; ++e) acc[e] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = min(max(2 * yy - 1 + a, 0), Ho - 1);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int ox = min(max(2 * xx - 1 + bb, 0), Wo - 1);
        Grp<T, G> g;
        g.load(b + ((long long)oy * Wo + ox) * C);
        const float wgt = wt[a] * wt[bb];
#pragma unroll
        for (int e = 0; e < G; ++e) acc[e] = fmaf(wgt, g.v[e], acc[e]);
      }
    }
    Grp<T, G> o;
#pragma unroll
    for (int e = 0; e < G; ++e) o.v[e] = acc[e];
    o.store(dx + (long long)i * G);
  }
}

// out = ca*a + cb*b with ca = a0 + a1*alpha, cb = b0 + b1*alpha (b may be null)
template <typename T>
__global__ void __launch_bounds__(256)
axpby_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, long long n,
             float a0, float a1, float b0, float b1, const float *__restrict__ alpha_dev) {
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
    /* 32-bit index arithmetic whenever it fits: the 64-bit div/mod chain per element made   \
       these kernels ALU-bound */                                                              \
    const bool small = work * 4 < (1ll << 31);                                                 \
    if (C % 8 == 0) {                                                                          \
      const int grid = bw_grid(work / 8, 256);                                                 \
      PG_DISPATCH_DTYPE(dtype, T, {                                                            \
        if (small) kernel<T, 8, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(             \
                       (const T *)x, (T *)y, N, H, W, C);                                      \
        else kernel<T, 8, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(                  \
                 (const T *)x, (T *)y, N, H, W, C);                                            \
      });                                                                                      \
    } else {                                                                                   \
      const int grid = bw_grid(work, 256);                                                     \
      PG_DISPATCH_DTYPE(dtype, T, {                                                            \
        if (small) kernel<T, 1, unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(             \
                       (const T *)x, (T *)y, N, H, W, C);                                      \
        else kernel<T, 1, long long><<<grid, 256, 0, (cudaStream_t)stream>>>(                  \
                 (const T *)x, (T *)y, N, H, W, C);                                            \
      });                                                                                      \
    }                                                                                          \
    PG_CHECK_LAUNCH(#fn);                                                                      \
  }

PG_RESAMPLE_ENTRY(pg_avgpool2, avgpool2_kernel, (long long)N * (H / 2) * (W / 2) * C, 1)
PG_RESAMPLE_ENTRY(pg_avgpool2_bwd, avgpool2_bwd_kernel, (long long)N * H * W * C, 1)
PG_RESAMPLE_ENTRY(pg_upsample2, upsample2_kernel, (long long)N * 4 * H * W * C, 0)
PG_RESAMPLE_ENTRY(pg_upsample2_bwd, upsample2_bwd_kernel, (long long)N * H * W * C, 0)

extern "C" int pg_blend(const void *a, const void *b, void *out, long long n,
                        const float *alpha_dev, int dtype, void *stream) {
  PG_CHECK_ARG(a && b && out && alpha_dev, "pg_blend: null pointer");
  PG_CHECK_ARG(n > 0, "pg_blend: n must be > 0");
  const int grid = bw_grid((n + 7) / 8, 256);
  PG_DISPATCH_DTYPE(dtype, T, axpby_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
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
  PG_DISPATCH_DTYPE(dtype, T, axpby_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                                  (const T *)x, (const T *)nullptr, (T *)out, n, c0, c1, 0.f,
                                  0.f, alpha_dev));
  PG_CHECK_LAUNCH("pg_scale");
}

extern "C" int pg_tanh_fwd(const float *x, float *y, long long n, void *stream) {
  PG_CHECK_ARG(x && y && n > 0, "pg_tanh_fwd: bad args");
  tanh_kernel<<<bw_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, nullptr, y, n, 0);
  PG_CHECK_LAUNCH("pg_tanh_fwd");
}

extern "C" int pg_tanh_bwd(const float *dy, const float *y, float *dx, long long n,
                           void *stream) {
  PG_CHECK_ARG(dy && y && dx && n > 0, "pg_tanh_bwd: bad args");
  tanh_kernel<<<bw_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, 1);
  PG_CHECK_LAUNCH("pg_tanh_bwd");
}
