// 1x1 heads with a 3- (or 1-, 4-) channel image side: from_rgb, to_rgb, final linear.
// Reference: EqualConv2d(3, C, 1) / EqualConv2d(C, 3, 1) / EqualLinear(C, 1) at
// progan_modules.py:195-200, 270-276, 280 (aten::convolution / addmm with K or N = 3,
// i.e. HBM-bound, never a tensor-core op — SURVEY.md §8 a10,a11,a13).
//
// img side: NCHW fp32 [N,K,HW] (exactly what the train scripts pass / receive);
// act side: NHWC [N*HW, C].  Logical weight w(c,k) = w[c*w_sc + k*w_sk].
#include "common.cuh"

namespace pg {

constexpr int kMaxK = 4;

// I = index type of the pixel arithmetic: unsigned (one 32-bit division per element; the 64-bit
// divisions of the generic form made these kernels ALU-bound at ~25% of HBM bandwidth) or
// long long for tensors with more than 2^31 elements.
template <typename T, typename I>
__global__ void __launch_bounds__(256)
pw_expand_kernel(const float *__restrict__ img, const float *__restrict__ w,
                 const float *__restrict__ bias, T *__restrict__ act, int N, I HW,
                 int K, int C, int w_sc, int w_sk, float scale) {
  extern __shared__ float sw[];  // [K][C] then bias[C]
  float *sb = sw + K * C;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    sw[i] = w[(long long)c * w_sc + (long long)k * w_sk] * scale;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) sb[c] = bias ? bias[c] : 0.f;
  __syncthreads();
  const I nch = (I)(C >> 3);
  const I total = (I)N * HW * nch;
  constexpr int UN = 4;                       // independent elements per thread: bytes in flight
  const I stride = (I)gridDim.x * (I)blockDim.x;
  // blockDim (256) is a multiple of nch (C/8 <= 256, power of two in this network), so a thread
  // always works on the same 8-channel group: its weights and bias live in registers and the
  // loop body has no shared-memory reads at all
  const bool fixed_group = (256 % (int)nch) == 0;
  const int cj_fixed = (int)(((I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x) % nch);
  float wr[kMaxK][8], br[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    br[e] = sb[cj_fixed * 8 + e];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) wr[k][e] = (k < K) ? sw[k * C + cj_fixed * 8 + e] : 0.f;
  }
  for (I i0 = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; i0 < total; i0 += stride * UN) {
    float xin[UN][kMaxK];
    I pixs[UN];
    int cjs[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const I i = i0 + (I)u * stride;
      const I pix = i / nch;
      pixs[u] = pix;
      cjs[u] = (int)(i - pix * nch);
      const I n = pix / HW, hw = pix - n * HW;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        xin[u][k] = (k < K && i < total) ? img[((long long)n * K + k) * (long long)HW + (long long)hw] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (i0 + (I)u * stride >= total) break;
      F8 o;
      if (fixed_group) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = br[e];
#pragma unroll
          for (int k = 0; k < kMaxK; ++k) v = fmaf(xin[u][k], wr[k][e], v);
          o.v[e] = v;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cjs[u] * 8 + e;
          float v = sb[c];
#pragma unroll
          for (int k = 0; k < kMaxK; ++k)
            if (k < K) v = fmaf(xin[u][k], sw[k * C + c], v);
          o.v[e] = v;
        }
      }
      st8(act + (long long)pixs[u] * C + (long long)cjs[u] * 8, o);
    }
  }
}

// CPL = 16-byte chunks per lane (C <= 8*TPP*CPL): chunk 0 multiplies weights held in registers,
// further chunks (C > 256) read theirs from shared memory; every load of the UN pixels is
// issued before the arithmetic.
template <typename T, int TPP, typename I, int CPL>
__global__ void __launch_bounds__(256)
pw_reduce_kernel(const T *__restrict__ act, const float *__restrict__ w,
                 const float *__restrict__ bias, float *__restrict__ img, int N, I HW,
                 int K, int C, int w_sc, int w_sk, float scale) {
  extern __shared__ float sw[];  // [K][C]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    sw[i] = w[(long long)c * w_sc + (long long)k * w_sk] * scale;
  }
  __syncthreads();
  const int sub = threadIdx.x % TPP;
  const I ppb = (I)(blockDim.x / TPP);
  const int nch = C >> 3;
  const I P = (I)N * HW;
  const I Pr = ((P + ppb - 1) / ppb) * ppb;
  constexpr int UN = 4;                       // pixels per sub-warp in flight
  const I pstride = (I)gridDim.x * ppb;
  // one chunk per lane (C <= 8*TPP): the lane's weights stay in registers
  float wr[kMaxK][8];
#pragma unroll
  for (int k = 0; k < kMaxK; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[k][e] = (k < K && sub < nch) ? sw[k * C + sub * 8 + e] : 0.f;
  for (I pix0 = (I)blockIdx.x * ppb + (I)(threadIdx.x / TPP); pix0 < Pr; pix0 += pstride * UN) {
    float acc[UN][kMaxK];
    typename RawOf<T>::type raw[UN][CPL];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const I pix = pix0 + (I)u * pstride;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int ch = sub + i * TPP;
        if (pix < P && ch < nch) raw[u][i] = ldraw8(act + (long long)pix * C + (long long)ch * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const I pix = pix0 + (I)u * pstride;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k) acc[u][k] = 0.f;
      if (pix < P) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int ch = sub + i * TPP;
          if (ch < nch) {
            const F8 v = unpack8(raw[u][i]);
            if (i == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int k = 0; k < kMaxK; ++k) acc[u][k] = fmaf(v.v[e], wr[k][e], acc[u][k]);
            } else {
#pragma unroll
              for (int k = 0; k < kMaxK; ++k)
                if (k < K) {
                  const float4 w0 = *reinterpret_cast<const float4 *>(sw + k * C + ch * 8);
                  const float4 w1 = *reinterpret_cast<const float4 *>(sw + k * C + ch * 8 + 4);
                  float a_ = acc[u][k];
                  a_ = fmaf(v.v[0], w0.x, a_); a_ = fmaf(v.v[1], w0.y, a_);
                  a_ = fmaf(v.v[2], w0.z, a_); a_ = fmaf(v.v[3], w0.w, a_);
                  a_ = fmaf(v.v[4], w1.x, a_); a_ = fmaf(v.v[5], w1.y, a_);
                  a_ = fmaf(v.v[6], w1.z, a_); a_ = fmaf(v.v[7], w1.w, a_);
                  acc[u][k] = a_;
                }
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const I pix = pix0 + (I)u * pstride;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k) {
#pragma unroll
        for (int o = TPP / 2; o > 0; o >>= 1) acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], o);
      }
      if (pix < P && sub == 0) {
        const I n = pix / HW, hw = pix - n * HW;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k)
          if (k < K) img[((long long)n * K + k) * (long long)HW + (long long)hw] = acc[u][k] + (bias ? bias[k] : 0.f);
      }
    }
  }
}

// dw(c,k) += scale * sum_pix act[pix,c] * img[k,pix]
template <typename T, typename I>
__global__ void __launch_bounds__(256)
pw_wgrad_kernel(const T *__restrict__ act, const float *__restrict__ img,
                float *__restrict__ dw, int N, I HW, int K, int C, int w_sc, int w_sk,
                float scale) {
  extern __shared__ float sm[];  // [rows][K][C]
  const int nch = C >> 3;
  const int rows = blockDim.x / nch;
  const int cj = threadIdx.x % nch, rj = threadIdx.x / nch;
  const I P = (I)N * HW;
  float acc[kMaxK][8];
#pragma unroll
  for (int k = 0; k < kMaxK; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  if (rj < rows) {
    constexpr int UN = 4;                     // pixels in flight per thread
    const I pstride = (I)gridDim.x * (I)rows;
    for (I pix0 = (I)blockIdx.x * (I)rows + (I)rj; pix0 < P; pix0 += pstride * UN) {
      typename RawOf<T>::type raw[UN];
      float g[UN][kMaxK];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const I pix = pix0 + (I)u * pstride;
        if (pix < P) {
          const I n = pix / HW, hw = pix - n * HW;
          raw[u] = ldraw8(act + (long long)pix * C + (long long)cj * 8);
#pragma unroll
          for (int k = 0; k < kMaxK; ++k)
            g[u][k] = (k < K) ? img[((long long)n * K + k) * (long long)HW + (long long)hw] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const I pix = pix0 + (I)u * pstride;
        if (pix < P) {
          const F8 v = unpack8(raw[u]);
#pragma unroll
          for (int k = 0; k < kMaxK; ++k) {
            if (k < K) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[k][e] = fmaf(v.v[e], g[u][k], acc[k][e]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sm[(rj * K + k) * C + cj * 8 + e] = acc[k][e];
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += sm[(r * K + k) * C + c];
    atomicAdd(dw + (long long)c * w_sc + (long long)k * w_sk, s * scale);
  }
}

__global__ void __launch_bounds__(256)
img_chansum_kernel(const float *__restrict__ img, float *__restrict__ out, int N, long long HW,
                   int K) {
  __shared__ float red[32];
  // grid.y = plane (n*K + k); grid.x strides over HW
  const int plane = blockIdx.y;
  const int k = plane % K;
  const float *p = img + (long long)plane * HW;
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
       i += (long long)gridDim.x * blockDim.x)
    s += p[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out + k, s);
}

}  // namespace pg

using namespace pg;

static int check_pw(const char *name, int N, long long HW, int K, int C) {
  PG_CHECK_ARG(N > 0 && HW > 0, "%s: bad dims", name);
  PG_CHECK_ARG(K >= 1 && K <= kMaxK, "%s: image channels K=%d must be in 1..%d", name, K, kMaxK);
  PG_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 2048, "%s: need C %% 8 == 0 and C <= 2048 (C=%d)",
               name, C);
  return PG_OK;
}

extern "C" int pg_pw_expand(const float *img, const float *w, const float *bias, void *act, int N,
                            long long HW, int K, int C, int w_sc, int w_sk, float scale,
                            int dtype, void *stream) {
  PG_CHECK_ARG(img && w && act, "pg_pw_expand: null pointer");
  if (int rc = check_pw("pg_pw_expand", N, HW, K, C)) return rc;
  const long long total = (long long)N * HW * (C / 8);
  const int grid = bw_grid(total, 256);
  const size_t smem = (size_t)(K + 1) * C * sizeof(float);
  const bool small = total + (long long)grid * 256 < (1ll << 31);
  PG_DISPATCH_DTYPE(dtype, T, {
    if (small)
      pw_expand_kernel<T, unsigned><<<grid, 256, smem, (cudaStream_t)stream>>>(
          img, w, bias, (T *)act, N, (unsigned)HW, K, C, w_sc, w_sk, scale);
    else
      pw_expand_kernel<T, long long><<<grid, 256, smem, (cudaStream_t)stream>>>(
          img, w, bias, (T *)act, N, HW, K, C, w_sc, w_sk, scale);
  });
  PG_CHECK_LAUNCH("pg_pw_expand");
}

extern "C" int pg_pw_reduce(const void *act, const float *w, const float *bias, float *img, int N,
                            long long HW, int K, int C, int w_sc, int w_sk, float scale,
                            int dtype, void *stream) {
  PG_CHECK_ARG(img && w && act, "pg_pw_reduce: null pointer");
  if (int rc = check_pw("pg_pw_reduce", N, HW, K, C)) return rc;
  const int nch = C / 8;
  const size_t smem = (size_t)K * C * sizeof(float);
  const long long P = (long long)N * HW;
  cudaStream_t s = (cudaStream_t)stream;
#define PG_LAUNCH_PWR(TPP, CPL)                                                             \
  {                                                                                         \
    const int grid = bw_grid(P, 256 / TPP);                                                 \
    if (P + (long long)grid * 256 < (1ll << 31))                                            \
      pw_reduce_kernel<T, TPP, unsigned, CPL><<<grid, 256, smem, s>>>(                      \
          (const T *)act, w, bias, img, N, (unsigned)HW, K, C, w_sc, w_sk, scale);          \
    else                                                                                    \
      pw_reduce_kernel<T, TPP, long long, CPL><<<grid, 256, smem, s>>>(                     \
          (const T *)act, w, bias, img, N, HW, K, C, w_sc, w_sk, scale);                    \
  }
  PG_DISPATCH_DTYPE(dtype, T, {
    if (nch <= 4) PG_LAUNCH_PWR(4, 1)
    else if (nch <= 8) PG_LAUNCH_PWR(8, 1)
    else if (nch <= 16) PG_LAUNCH_PWR(16, 1)
    else if (nch <= 32) PG_LAUNCH_PWR(32, 1)
    else if (nch <= 64) PG_LAUNCH_PWR(32, 2)
    else if (nch <= 128) PG_LAUNCH_PWR(32, 4)
    else PG_LAUNCH_PWR(32, 8)
  });
#undef PG_LAUNCH_PWR
  PG_CHECK_LAUNCH("pg_pw_reduce");
}

extern "C" int pg_pw_wgrad(const void *act, const float *img, float *dw, int N, long long HW,
                           int K, int C, int w_sc, int w_sk, float scale, int dtype,
                           void *stream) {
  PG_CHECK_ARG(img && dw && act, "pg_pw_wgrad: null pointer");
  if (int rc = check_pw("pg_pw_wgrad", N, HW, K, C)) return rc;
  const int nch = C / 8;
  const int rows = 256 / nch;
  PG_CHECK_ARG(rows >= 1, "pg_pw_wgrad: C too large");
  const long long P = (long long)N * HW;
  const int grid = bw_grid(P, rows * 16, 4);
  const size_t smem = (size_t)rows * K * C * sizeof(float);
  PG_DISPATCH_DTYPE(dtype, T, {
    if (P + (long long)grid * 256 < (1ll << 31))
      pw_wgrad_kernel<T, unsigned><<<grid, 256, smem, (cudaStream_t)stream>>>(
          (const T *)act, img, dw, N, (unsigned)HW, K, C, w_sc, w_sk, scale);
    else
      pw_wgrad_kernel<T, long long><<<grid, 256, smem, (cudaStream_t)stream>>>(
          (const T *)act, img, dw, N, HW, K, C, w_sc, w_sk, scale);
  });
  PG_CHECK_LAUNCH("pg_pw_wgrad");
}

extern "C" int pg_img_chansum(const float *img, float *out, int N, long long HW, int K,
                              void *stream) {
  PG_CHECK_ARG(img && out, "pg_img_chansum: null pointer");
  PG_CHECK_ARG(N > 0 && HW > 0 && K > 0 && (long long)N * K <= 65535, "pg_img_chansum: bad dims");
  int gx = (int)((HW + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, N * K);
  img_chansum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, out, N, HW, K);
  PG_CHECK_LAUNCH("pg_img_chansum");
}
