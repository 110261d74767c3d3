// PixelNorm + LeakyReLU first- and second-order derivative kernels, bias-gradient
// column sum.  Reference ops: PixelNorm.forward (progan_modules.py:54-60) and
// nn.LeakyReLU(0.2) (:138,142), whose backward / double-backward the reference gets
// from autograd as chains of unfused mul/div/pow/sum kernels (SURVEY.md §8 a5,a6).
//
// Bandwidth-bound: a sub-warp of C/8 lanes (<=32) owns one pixel, every lane moves
// 16-byte vectors, channel reductions are xor-shuffles inside the sub-warp.
#include "common.cuh"

namespace pg {

template <int TPP>
__device__ __forceinline__ float subwarp_sum(float v) {
#pragma unroll
  for (int o = TPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MAXI = chunks of 8 channels per lane (C <= 8*TPP*MAXI); U = pixels per sub-warp per
// iteration whose loads are all issued before any arithmetic (bytes in flight).
// pool_w > 0: dy is the gradient of the 2x2-average-pooled activation ([N,H/2,W/2,C], the
// reference's bilinear x0.5, progan_modules.py:299) and is expanded on the fly (x 1/4), which
// fuses avgpool2_bwd into this kernel.  colsum != nullptr (first order only): the per-channel
// sum of the produced da — the bias gradient of the conv in front — is accumulated too.
template <typename T, int TPP, int MAXI, bool SECOND, int U, bool ADD>
__global__ void __launch_bounds__(256, 3)
pn_lrelu_grad_kernel(const T *__restrict__ t_in, const T *__restrict__ dy,
                     const T *__restrict__ y, const float *__restrict__ rr,
                     T *__restrict__ out0, T *__restrict__ out1, long long P, int C,
                     float slope, int use_pn, int pool_h, int pool_w,
                     float *__restrict__ colsum) {
  pg::grid_dep_sync();
  // first order : out0 = da (+ t_in when ADD: a second gradient contribution to the same
  //               pre-activation, summed here instead of by a separate add kernel)
  // second order: out0 = cot_dy, out1 = cot_a
  using Raw = typename RawOf<T>::type;
  const int sub = threadIdx.x % TPP;
  const long long ppb = blockDim.x / TPP;
  const int nch = C >> 3;
  const float invC = 1.f / (float)C;
  const float inv_slope = 1.f / slope;
  const long long stride = (long long)gridDim.x * ppb;
  float csum[MAXI][8];
#pragma unroll
  for (int i = 0; i < MAXI; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) csum[i][e] = 0.f;
  const long long Pr = ((P + ppb - 1) / ppb) * ppb;
  for (long long pix0 = blockIdx.x * ppb + threadIdx.x / TPP; pix0 < Pr; pix0 += stride * U) {
    Raw ry[U][MAXI], rd[U][MAXI], rt[U][MAXI];
    float rv[U];
    // ---- phase 1: issue every load of the U pixels
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pix = pix0 + u * stride;
      const bool live = pix < P;
      rv[u] = 1.f;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int ch = sub + i * TPP;
        if (live && ch < nch) {
          const long long off = pix * C + (long long)ch * 8;
          ry[u][i] = ldraw8(y + off);
          if (pool_w > 0) {
            const unsigned upix = (unsigned)pix;
            const unsigned lw = (unsigned)pool_w >> 16, lh = (unsigned)pool_h >> 16;
            const unsigned W_ = (unsigned)pool_w & 0xFFFFu, H_ = (unsigned)pool_h & 0xFFFFu;
            unsigned wq, hq, nq;
            if (lw) { wq = upix & (W_ - 1); hq = (upix >> (lw - 1)) & (H_ - 1); nq = upix >> (lw + lh - 2); }
            else { wq = upix % W_; hq = (upix / W_) % H_; nq = upix / (W_ * H_); }
            const long long pp = ((long long)nq * (H_ >> 1) + (hq >> 1)) * (W_ >> 1) + (wq >> 1);
            rd[u][i] = ldraw8(dy + pp * C + (long long)ch * 8);
          } else {
            rd[u][i] = ldraw8(dy + off);
          }
          if (SECOND || ADD) rt[u][i] = ldraw8(t_in + off);
        }
      }
      if (use_pn && live) rv[u] = rr[pix];
    }
    // ---- phase 2: arithmetic + stores, one pixel at a time
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pix = pix0 + u * stride;
      const bool live = pix < P;
      const float dscale = pool_w > 0 ? 0.25f : 1.f;
      F8 pv[MAXI], uv[MAXI], tv[MAXI];
      float mk[MAXI][8];
      float s_pu = 0.f, s_pt = 0.f, s_tu = 0.f;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int ch = sub + i * TPP;
        if (live && ch < nch) {
          const F8 yv = unpack8(ry[u][i]);
          const F8 dv = unpack8(rd[u][i]);
          if (SECOND || ADD) tv[i] = unpack8(rt[u][i]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const bool pos = yv.v[e] > 0.f;
            const float m = pos ? 1.f : slope;
            const float p = pos ? yv.v[e] : yv.v[e] * inv_slope;
            mk[i][e] = m;
            pv[i].v[e] = p;
            uv[i].v[e] = m * dv.v[e] * dscale;
            s_pu += p * uv[i].v[e];
            if (SECOND) {
              s_pt += p * tv[i].v[e];
              s_tu += tv[i].v[e] * uv[i].v[e];
            }
          }
        }
      }
      const float r = rv[u];
      if (use_pn) {
        s_pu = subwarp_sum<TPP>(s_pu);
        if (SECOND) {
          s_pt = subwarp_sum<TPP>(s_pt);
          s_tu = subwarp_sum<TPP>(s_tu);
        }
      }
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int ch = sub + i * TPP;
        if (live && ch < nch) {
          const long long off = pix * C + (long long)ch * 8;
          F8 o0, o1;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float p = pv[i].v[e], uu = uv[i].v[e];
            if (!SECOND) {
              o0.v[e] = use_pn ? r * (uu - p * s_pu * invC) : uu;
              if (ADD) o0.v[e] += tv[i].v[e];
              csum[i][e] += o0.v[e];
            } else {
              const float t = tv[i].v[e];
              if (use_pn) {
                o0.v[e] = mk[i][e] * r * (t - p * s_pt * invC);
                o1.v[e] = r * r * invC *
                          (3.f * invC * s_pt * s_pu * p - s_tu * p - s_pu * t - s_pt * uu);
              } else {
                o0.v[e] = mk[i][e] * t;
                o1.v[e] = 0.f;
              }
            }
          }
          st8(out0 + off, o0);
          if (SECOND) st8(out1 + off, o1);
        }
      }
    }
  }
  if (!SECOND && colsum != nullptr) {
    // block reduction over the pixel slots, then one atomic per channel per block
    extern __shared__ float cs_sm[];              // [ppb][C]
    const int slot = threadIdx.x / TPP;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int ch = sub + i * TPP;
      if (ch < nch) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cs_sm[slot * C + ch * 8 + e] = csum[i][e];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
      for (int sl = 0; sl < (int)ppb; ++sl) a += cs_sm[sl * C + c];
      atomicAdd(colsum + c, a);
    }
  }
}

// First-order backward through PixelNorm + LeakyReLU + 2x2 average pool, one sub-warp per POOLED
// pixel: the four full-resolution pixels of the quad share one dy vector, which is loaded once
// (the generic kernel re-reads it from four different CTAs and pays the index arithmetic four
// times).  C <= 8*TPP.
template <typename T, int TPP, bool ADD>
__global__ void __launch_bounds__(256, 3)
pn_lrelu_bwd_pooled_kernel(const T *__restrict__ addend, const T *__restrict__ dy,
                           const T *__restrict__ y, const float *__restrict__ rr,
                           T *__restrict__ da, unsigned NQ, unsigned H2, unsigned W2, int C,
                           float slope, int use_pn, float *__restrict__ colsum) {
  pg::grid_dep_sync();
  using Raw = typename RawOf<T>::type;
  const int sub = threadIdx.x % TPP;
  const unsigned qpb = blockDim.x / TPP;
  const int nch = C >> 3;
  const bool lane_live = sub < nch;
  const float invC = 1.f / (float)C;
  const float inv_slope = 1.f / slope;
  const unsigned W = 2u * W2;
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const unsigned NQr = ((NQ + qpb - 1) / qpb) * qpb;
  for (unsigned q = blockIdx.x * qpb + threadIdx.x / TPP; q < NQr; q += gridDim.x * qpb) {
    const bool live = q < NQ && lane_live;
    const unsigned t = q / W2, w2 = q - t * W2;
    const unsigned n = t / H2, h2 = t - n * H2;
    const long long p00 = ((long long)n * (2u * H2) + 2u * h2) * W + 2u * w2;
    const long long pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
    Raw ry[4], ra[4], rd;
    float rv[4] = {1.f, 1.f, 1.f, 1.f};
    if (live) {
      rd = ldraw8(dy + (long long)q * C + (long long)sub * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ry[u] = ldraw8(y + pix[u] * C + (long long)sub * 8);
        if (ADD) ra[u] = ldraw8(addend + pix[u] * C + (long long)sub * 8);
      }
    }
    if (use_pn && q < NQ) {
#pragma unroll
      for (int u = 0; u < 4; ++u) rv[u] = rr[pix[u]];
    }
    F8 dv;
    if (live) dv = unpack8(rd);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      F8 pv, uv;
      float s_pu = 0.f;
      if (live) {
        const F8 yv = unpack8(ry[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool pos = yv.v[e] > 0.f;
          pv.v[e] = pos ? yv.v[e] : yv.v[e] * inv_slope;
          uv.v[e] = (pos ? 0.25f : 0.25f * slope) * dv.v[e];
          s_pu += pv.v[e] * uv.v[e];
        }
      }
      if (use_pn) s_pu = subwarp_sum<TPP>(s_pu);
      if (live) {
        F8 o;
        F8 av;
        if (ADD) av = unpack8(ra[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o.v[e] = use_pn ? rv[u] * (uv.v[e] - pv.v[e] * s_pu * invC) : uv.v[e];
          if (ADD) o.v[e] += av.v[e];
          csum[e] += o.v[e];
        }
        st8(da + pix[u] * C + (long long)sub * 8, o);
      }
    }
  }
  if (colsum != nullptr) {
    extern __shared__ float cs_sm[];              // [qpb][C]
    const int slot = threadIdx.x / TPP;
    if (lane_live) {
#pragma unroll
      for (int e = 0; e < 8; ++e) cs_sm[slot * C + sub * 8 + e] = csum[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
      for (int sl = 0; sl < (int)qpb; ++sl) a += cs_sm[sl * C + c];
      atomicAdd(colsum + c, a);
    }
  }
}

template <typename T, bool SECOND>
static int launch_pn_grad(const T *t, const T *dy, const T *y, const float *r, T *o0, T *o1,
                          long long P, int C, float slope, int use_pn, int pool_h, int pool_w,
                          float *colsum, cudaStream_t s) {
  const int nch = C / 8;
  if (!SECOND && pool_w > 0 && nch <= 32 && P < (1ll << 31)) {
    // quad kernel (pool_h/pool_w still carry the log2 encoding in their high halves)
    const unsigned H_ = (unsigned)pool_h & 0xFFFFu, W_ = (unsigned)pool_w & 0xFFFFu;
    const unsigned NQ = (unsigned)(P / 4);
#define PG_LAUNCH_PQ(TPP)                                                                  \
  {                                                                                        \
    const int qpb = 256 / TPP;                                                             \
    const int grid = bw_grid(NQ, qpb, 3);                                                  \
    const size_t sm = colsum ? (size_t)qpb * C * sizeof(float) : 0;                        \
    if (t != nullptr)                                                                      \
      pg::launcher(pn_lrelu_bwd_pooled_kernel<T, TPP, true>, grid, 256, sm, s)(                      \
          t, dy, y, r, o0, NQ, H_ / 2, W_ / 2, C, slope, use_pn, colsum);                  \
    else                                                                                   \
      pg::launcher(pn_lrelu_bwd_pooled_kernel<T, TPP, false>, grid, 256, sm, s)(                     \
          t, dy, y, r, o0, NQ, H_ / 2, W_ / 2, C, slope, use_pn, colsum);                  \
  }
    if (nch <= 4) PG_LAUNCH_PQ(4)
    else if (nch <= 8) PG_LAUNCH_PQ(8)
    else if (nch <= 16) PG_LAUNCH_PQ(16)
    else PG_LAUNCH_PQ(32)
#undef PG_LAUNCH_PQ
    return PG_OK;
  }
#define PG_LAUNCH_PN(TPP, MAXI)                                                           \
  {                                                                                       \
    const long long ppb = 256 / TPP;                                                      \
    constexpr int U_ = (MAXI == 1) ? (SECOND ? 2 : 4) : 1;                                 \
    const int grid = bw_grid(P, (int)ppb * U_);                                           \
    const size_t sm = (!SECOND && colsum) ? (size_t)ppb * C * sizeof(float) : 0;          \
    if (!SECOND && t != nullptr)                                                          \
      pg::launcher(pn_lrelu_grad_kernel<T, TPP, MAXI, false, U_, true>, grid, 256, sm, s)(          \
          t, dy, y, r, o0, o1, P, C, slope, use_pn, pool_h, pool_w, colsum);              \
    else                                                                                  \
      pg::launcher(pn_lrelu_grad_kernel<T, TPP, MAXI, SECOND, U_, false>, grid, 256, sm, s)(        \
          t, dy, y, r, o0, o1, P, C, slope, use_pn, pool_h, pool_w, colsum);              \
  }
  if (nch <= 4) PG_LAUNCH_PN(4, 1)
  else if (nch <= 8) PG_LAUNCH_PN(8, 1)
  else if (nch <= 16) PG_LAUNCH_PN(16, 1)
  else if (nch <= 32) PG_LAUNCH_PN(32, 1)
  else if (nch <= 64) PG_LAUNCH_PN(32, 2)
  else if (nch <= 128) PG_LAUNCH_PN(32, 4)
  else {
    set_error("pn_lrelu: C=%d > 1024 unsupported", C);
    return PG_ERR_UNSUPPORTED;
  }
#undef PG_LAUNCH_PN
  return PG_OK;
}

// out[c] += sum_pix x[pix, c]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T *__restrict__ x, float *__restrict__ out, long long P, int C) {
  pg::grid_dep_sync();
  extern __shared__ float sm[];  // [rows][C]
  const int nch = C >> 3;
  const int rows = blockDim.x / nch;
  const int cj = threadIdx.x % nch, rj = threadIdx.x / nch;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rj < rows) {
    using Raw = typename RawOf<T>::type;
    const long long stride = (long long)gridDim.x * rows;
    long long pix = (long long)blockIdx.x * rows + rj;
    for (; pix + 3 * stride < P; pix += 4 * stride) {      // four independent loads in flight
      Raw raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = ldraw8(x + (pix + u * stride) * C + (long long)cj * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const F8 v = unpack8(raw[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v.v[e];
      }
    }
    for (; pix < P; pix += stride) {
      F8 v = ld8(x + pix * C + (long long)cj * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v.v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[rj * C + cj * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += sm[r * C + c];
    atomicAdd(out + c, s);
  }
}

}  // namespace pg

using namespace pg;

// power-of-two H and W: put (log2 + 1) into the high 16 bits so the kernel can shift/mask
static void pg_encode_pool(int *h, int *w) {
  auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return ((1 << l) == v) ? l + 1 : 0; };
  const int lh = lg(*h), lw = lg(*w);
  if (*w > 0 && lh && lw) {
    *h |= lh << 16;
    *w |= lw << 16;
  }
}

// Stand-alone forward y = lrelu(a * rsqrt(mean_c a^2 + 1e-8)) with the per-pixel statistic r
// written for the backward kernels: the activation of layers too wide for the conv epilogue
// (more than 256 output channels = several N tiles).  One warp per pixel, MAXV 16-byte vectors
// per lane kept in registers between the reduction and the store (C <= 256 * MAXV).
template <typename T, int MAXV>
__global__ void __launch_bounds__(256)
pn_lrelu_fwd_kernel(const T *__restrict__ a, T *__restrict__ y, float *__restrict__ r, long long P,
                    int C, float slope, int use_pn) {
  pg::grid_dep_sync();
  using Raw = typename RawOf<T>::type;
  const int lane = threadIdx.x & 31;
  const int nch = C >> 3;
  const long long wpb = blockDim.x >> 5;
  for (long long pix = blockIdx.x * wpb + (threadIdx.x >> 5); pix < P; pix += gridDim.x * wpb) {
    Raw raw[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int ch = lane + i * 32;
      if (ch < nch) raw[i] = ldraw8(a + pix * C + (long long)ch * 8);
    }
    F8 v[MAXV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nch) {
        v[i] = unpack8(raw[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(v[i].v[e], v[i].v[e], ss);
      }
    }
    float rs = 1.f;
    if (use_pn) {
      rs = rsqrtf(warp_sum(ss) / (float)C + 1e-8f);
      if (lane == 0) r[pix] = rs;
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int ch = lane + i * 32;
      if (ch < nch) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float pv = v[i].v[e] * rs;
          v[i].v[e] = pv > 0.f ? pv : pv * slope;
        }
        st8(y + pix * C + (long long)ch * 8, v[i]);
      }
    }
  }
}

extern "C" int pg_pn_lrelu_fwd(const void *a, void *y, float *r, long long P, int C, float slope,
                               int use_pn, int dtype, void *stream) {
  PG_CHECK_ARG(a && y && (r || !use_pn), "pg_pn_lrelu_fwd: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0 && C <= 1024, "pg_pn_lrelu_fwd: need C %% 8 == 0, C <= 1024 (C=%d)", C);
  const long long want = (P + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  PG_DISPATCH_DTYPE(dtype, T, {
    if (C <= 256)
      pg::launcher(pn_lrelu_fwd_kernel<T, 1>, grid, 256, 0, (cudaStream_t)stream)((const T *)a, (T *)y, r, P, C, slope, use_pn);
    else if (C <= 512)
      pg::launcher(pn_lrelu_fwd_kernel<T, 2>, grid, 256, 0, (cudaStream_t)stream)((const T *)a, (T *)y, r, P, C, slope, use_pn);
    else
      pg::launcher(pn_lrelu_fwd_kernel<T, 4>, grid, 256, 0, (cudaStream_t)stream)((const T *)a, (T *)y, r, P, C, slope, use_pn);
  });
  PG_CHECK_LAUNCH("pg_pn_lrelu_fwd");
}

extern "C" int pg_pn_lrelu_bwd(const void *dy, const void *y, const float *r, void *da,
                               long long P, int C, float slope, int use_pn, int pool_h,
                               int pool_w, float *colsum, const void *addend, int dtype,
                               void *stream) {
  PG_CHECK_ARG((pool_h == 0) == (pool_w == 0) && pool_h % 2 == 0 && pool_w % 2 == 0,
               "pg_pn_lrelu_bwd: pooled form needs even H and W");
  PG_CHECK_ARG(pool_w == 0 || P % ((long long)pool_h * pool_w) == 0, "pg_pn_lrelu_bwd: P is not N*H*W");
  PG_CHECK_ARG(pool_w == 0 || (P < (1ll << 31) && pool_h < 65536 && pool_w < 65536), "pg_pn_lrelu_bwd: too large");
  pg_encode_pool(&pool_h, &pool_w);
  PG_CHECK_ARG(dy && y && da && (r || !use_pn), "pg_pn_lrelu_bwd: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0, "pg_pn_lrelu_bwd: need C %% 8 == 0 (C=%d)", C);
  PG_CHECK_ARG(slope > 0.f, "pg_pn_lrelu_bwd: slope must be > 0");
  PG_DISPATCH_DTYPE(dtype, T, {
    int rc = launch_pn_grad<T, false>((const T *)addend, (const T *)dy, (const T *)y, r, (T *)da, nullptr,
                                      P, C, slope, use_pn, pool_h, pool_w, colsum,
                                      (cudaStream_t)stream);
    if (rc) return rc;
  });
  PG_CHECK_LAUNCH("pg_pn_lrelu_bwd");
}

extern "C" int pg_pn_lrelu_bwd_bwd(const void *t, const void *dy, const void *y, const float *r,
                                   void *cot_dy, void *cot_a, long long P, int C, float slope,
                                   int use_pn, int pool_h, int pool_w, int dtype, void *stream) {
  PG_CHECK_ARG((pool_h == 0) == (pool_w == 0) && pool_h % 2 == 0 && pool_w % 2 == 0,
               "pg_pn_lrelu_bwd_bwd: pooled form needs even H and W");
  PG_CHECK_ARG(pool_w == 0 || (P < (1ll << 31) && pool_h < 65536 && pool_w < 65536), "pg_pn_lrelu_bwd_bwd: too large");
  pg_encode_pool(&pool_h, &pool_w);
  PG_CHECK_ARG(t && dy && y && cot_dy && cot_a && (r || !use_pn),
               "pg_pn_lrelu_bwd_bwd: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0, "pg_pn_lrelu_bwd_bwd: need C %% 8 == 0 (C=%d)", C);
  PG_CHECK_ARG(slope > 0.f, "pg_pn_lrelu_bwd_bwd: slope must be > 0");
  PG_DISPATCH_DTYPE(dtype, T, {
    int rc = launch_pn_grad<T, true>((const T *)t, (const T *)dy, (const T *)y, r, (T *)cot_dy,
                                     (T *)cot_a, P, C, slope, use_pn, pool_h, pool_w, nullptr,
                                     (cudaStream_t)stream);
    if (rc) return rc;
  });
  PG_CHECK_LAUNCH("pg_pn_lrelu_bwd_bwd");
}

extern "C" int pg_colsum(const void *x, float *out, long long P, int C, int dtype, void *stream) {
  PG_CHECK_ARG(x && out, "pg_colsum: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0 && C <= 2048, "pg_colsum: need C %% 8 == 0, C<=2048");
  const int nch = C / 8;
  const int rows = 256 / nch;
  const int grid = bw_grid(P, rows * 16, 4);
  const size_t smem = (size_t)rows * C * sizeof(float);
  PG_DISPATCH_DTYPE(dtype, T, pg::launcher(colsum_kernel<T>, grid, 256, smem, (cudaStream_t)stream)(
                                  (const T *)x, out, P, C));
  PG_CHECK_LAUNCH("pg_colsum");
}
