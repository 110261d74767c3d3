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
pw_expand_generic_kernel(const float *__restrict__ img, const float *__restrict__ w,
                 const float *__restrict__ bias, T *__restrict__ act, int N, I HW,
                 int K, int C, int w_sc, int w_sk, float scale) {
  pg::grid_dep_sync();
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
pw_reduce_generic_kernel(const T *__restrict__ act, const float *__restrict__ w,
                 const float *__restrict__ bias, float *__restrict__ img, int N, I HW,
                 int K, int C, int w_sc, int w_sk, float scale) {
  pg::grid_dep_sync();
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
pw_wgrad_generic_kernel(const T *__restrict__ act, const float *__restrict__ img,
                float *__restrict__ dw, int N, I HW, int K, int C, int w_sc, int w_sk,
                float scale) {
  pg::grid_dep_sync();
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

// ------------------------------------------------------------------------------------------
// Warp-run forms (C = 8 * LPP * NCK channels, LPP = lanes per pixel in {1,2,4,8,16,32}; KT = number
// of image channels, compile time): a warp owns a run of 32 consecutive pixels.  On the image
// side lane l touches pixel l of the run (one coalesced 128-byte access per image plane); on the
// activation side the warp moves 512 contiguous bytes per instruction (32 / LPP pixels x LPP
// 16-byte chunks) and the per-pixel image values travel between the two mappings by warp
// shuffles.  These kernels are INSTRUCTION-bound, not latency-bound (ncu: issue slots 60-70 %
// busy at 0.3-0.4 of the HBM roofline in their first form), so everything per element that is not
// an FMA is removed: 32-bit index arithmetic, one division per lane per run, weights in
// registers, the image-channel count a template parameter, and loop trip counts that depend on
// kernel parameters only (a warp-uniform loop lets ptxas emit plain SHFL instead of
// WARPSYNC/ENDCOLLECTIVE sequences around every shuffle).
template <typename T, int LPP, int NCK, int KT>
__global__ void __launch_bounds__(256)
pw_expand_kernel(const float *__restrict__ img, const float *__restrict__ w,
                 const float *__restrict__ bias, T *__restrict__ act, int P, int HW,
                 int K, int w_sc, int w_sk, float scale) {
  pg::grid_dep_sync();
  constexpr int C = 8 * LPP * NCK;
  constexpr int PPW = 32 / LPP;                 // pixels per store instruction
  __shared__ float sw[NCK > 1 ? (KT + 1) * C : 1];
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPP, slot = lane / LPP;
  float wr[KT][8], br[8];
  if (NCK == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      br[e] = bias ? bias[sub * 8 + e] : 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        wr[k][e] = (k < K) ? w[(sub * 8 + e) * w_sc + k * w_sk] * scale : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < KT * C; i += blockDim.x) {
      const int k = i / C, c = i - k * C;
      sw[i] = (k < K) ? w[c * w_sc + k * w_sk] * scale : 0.f;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) sw[KT * C + c] = bias ? bias[c] : 0.f;
    __syncthreads();
  }
  const int runs = (P + 31) >> 5;
  const int wpb = blockDim.x >> 5;
  const int warp0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpb;
  constexpr int UN = 2;                          // runs in flight per warp
  const int n_iter = (runs + nwarps * UN - 1) / (nwarps * UN);     // warp-uniform trip count
  for (int it = 0; it < n_iter; ++it) {
    const int run0 = warp0 + it * nwarps * UN;
    float xk[UN][KT];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int pix = ((run0 + u * nwarps) << 5) + lane;
      const bool ok = pix < P;
      const int n = ok ? pix / HW : 0;
      const int hw = pix - n * HW;
      const float *src = img + (size_t)n * K * HW + hw;
#pragma unroll
      for (int k = 0; k < KT; ++k) xk[u][k] = (ok && k < K) ? src[(size_t)k * HW] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int base = (run0 + u * nwarps) << 5;
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const int pl = j * PPW + slot;
        float xs[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) xs[k] = __shfl_sync(0xffffffffu, xk[u][k], pl);
        const int pp = base + pl;
        if (pp < P) {
#pragma unroll
          for (int ck = 0; ck < NCK; ++ck) {
            const int cg = sub + ck * LPP;
            F8 o;
            if (NCK == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float v = br[e];
#pragma unroll
                for (int k = 0; k < KT; ++k) v = fmaf(xs[k], wr[k][e], v);
                o.v[e] = v;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float v = sw[KT * C + cg * 8 + e];
#pragma unroll
                for (int k = 0; k < KT; ++k) v = fmaf(xs[k], sw[k * C + cg * 8 + e], v);
                o.v[e] = v;
              }
            }
            st8(act + (size_t)pp * C + cg * 8, o);
          }
        }
      }
    }
  }
}

template <typename T, int LPP, int NCK, int KT>
__global__ void __launch_bounds__(256)
pw_reduce_kernel(const T *__restrict__ act, const float *__restrict__ w,
                 const float *__restrict__ bias, float *__restrict__ img, int P, int HW,
                 int K, int w_sc, int w_sk, float scale) {
  pg::grid_dep_sync();
  constexpr int C = 8 * LPP * NCK;
  constexpr int PPW = 32 / LPP;
  constexpr int RB = (LPP * NCK <= 8) ? LPP : (8 / NCK > 0 ? 8 / NCK : 1);   // rounds per load batch
  __shared__ float sw[NCK > 1 ? KT * C : 1];
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPP, slot = lane / LPP;
  float wr[KT][8];
  if (NCK == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int k = 0; k < KT; ++k)
        wr[k][e] = (k < K) ? w[(sub * 8 + e) * w_sc + k * w_sk] * scale : 0.f;
  } else {
    for (int i = threadIdx.x; i < KT * C; i += blockDim.x) {
      const int k = i / C, c = i - k * C;
      sw[i] = (k < K) ? w[c * w_sc + k * w_sk] * scale : 0.f;
    }
    __syncthreads();
  }
  float bk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) bk[k] = (bias && k < K) ? bias[k] : 0.f;
  const int runs = (P + 31) >> 5;
  const int wpb = blockDim.x >> 5;
  const int warp0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpb;
  const int n_iter = (runs + nwarps - 1) / nwarps;                 // warp-uniform trip count
  for (int it = 0; it < n_iter; ++it) {
    const int base = (warp0 + it * nwarps) << 5;
    float keep[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) keep[k] = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < LPP; j0 += RB) {
      typename RawOf<T>::type raw[RB][NCK];
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) {
        const int pp = base + (j0 + jj) * PPW + slot;
#pragma unroll
        for (int ck = 0; ck < NCK; ++ck)
          if (pp < P) raw[jj][ck] = ldraw8(act + (size_t)pp * C + (sub + ck * LPP) * 8);
      }
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) {
        const int j = j0 + jj;
        const int pp = base + j * PPW + slot;
        float acc[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = 0.f;
        if (pp < P) {
#pragma unroll
          for (int ck = 0; ck < NCK; ++ck) {
            const F8 v = unpack8(raw[jj][ck]);
            if (NCK == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int k = 0; k < KT; ++k) acc[k] = fmaf(v.v[e], wr[k][e], acc[k]);
            } else {
              const int cg = sub + ck * LPP;
#pragma unroll
              for (int k = 0; k < KT; ++k)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[k] = fmaf(v.v[e], sw[k * C + cg * 8 + e], acc[k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
#pragma unroll
          for (int o = LPP / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
          // pixel j*PPW + s of the run was summed by the lanes of slot s: hand it to lane j*PPW + s
          const float t = __shfl_sync(0xffffffffu, acc[k], (lane % PPW) * LPP);
          if (lane / PPW == j) keep[k] = t;
        }
      }
    }
    const int pix = base + lane;
    if (pix < P) {
      const int n = pix / HW, hw = pix - n * HW;
      float *dst = img + (size_t)n * K * HW + hw;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) dst[(size_t)k * HW] = keep[k] + bk[k];
    }
  }
}

// dw(c,k) += scale * sum_pix act[pix,c] * img[k,pix];  dbias(c) += sum_pix act[pix,c] (optional:
// the bias gradient of a from_rgb layer rides on the same pass over its output gradient)
template <typename T, int LPP, int KT>
__global__ void __launch_bounds__(256)
pw_wgrad_kernel(const T *__restrict__ act, const float *__restrict__ img, float *__restrict__ dw,
                float *__restrict__ dbias, int P, int HW, int K, int w_sc, int w_sk, float scale) {
  pg::grid_dep_sync();
  constexpr int C = 8 * LPP;
  constexpr int PPW = 32 / LPP;
  constexpr int RB = LPP <= 8 ? LPP : 8;
  __shared__ float sm[8][KT + 1][C];             // per-warp partial sums
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane % LPP, slot = lane / LPP;
  float acc[KT + 1][8];
#pragma unroll
  for (int k = 0; k <= KT; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  const int runs = (P + 31) >> 5;
  const int wpb = blockDim.x >> 5;
  const int warp0 = blockIdx.x * wpb + wid;
  const int nwarps = gridDim.x * wpb;
  const int n_iter = (runs + nwarps - 1) / nwarps;                 // warp-uniform trip count
  for (int it = 0; it < n_iter; ++it) {
    const int base = (warp0 + it * nwarps) << 5;
    const int pix = base + lane;
    const bool ok = pix < P;
    const int n = ok ? pix / HW : 0;
    const int hw = pix - n * HW;
    const float *src = img + (size_t)n * K * HW + hw;
    float gk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) gk[k] = (ok && k < K) ? src[(size_t)k * HW] : 0.f;
#pragma unroll
    for (int j0 = 0; j0 < LPP; j0 += RB) {
      typename RawOf<T>::type raw[RB];
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) {
        const int pp = base + (j0 + jj) * PPW + slot;
        if (pp < P) raw[jj] = ldraw8(act + (size_t)pp * C + sub * 8);
      }
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) {
        const int pl = (j0 + jj) * PPW + slot;
        float g[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) g[k] = __shfl_sync(0xffffffffu, gk[k], pl);
        if (base + pl < P) {
          const F8 v = unpack8(raw[jj]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int k = 0; k < KT; ++k) acc[k][e] = fmaf(v.v[e], g[k], acc[k][e]);
            acc[KT][e] += v.v[e];
          }
        }
      }
    }
  }
  // lanes with the same `sub` hold partial sums of the same channels
#pragma unroll
  for (int k = 0; k <= KT; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
      for (int o = LPP; o < 32; o <<= 1) acc[k][e] += __shfl_xor_sync(0xffffffffu, acc[k][e], o);
    }
  if (slot == 0) {
#pragma unroll
    for (int k = 0; k <= KT; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) sm[wid][k][sub * 8 + e] = acc[k][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (KT + 1) * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    if (k < K || (k == KT && dbias)) {
      float s = 0.f;
      for (int r = 0; r < wpb; ++r) s += sm[r][k][c];
      if (k < K) atomicAdd(dw + c * w_sc + k * w_sk, s * scale);
      else if (k == KT) atomicAdd(dbias + c, s);
    }
  }
}

__global__ void __launch_bounds__(256)
img_chansum_kernel(const float *__restrict__ img, float *__restrict__ out, int N, long long HW,
                   int K) {
  pg::grid_dep_sync();
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

// (LPP, NCK) of the warp-run kernels for C channels, or false when C is not served by them
static bool pw_shape(int C, int *lpp, int *nck) {
  const int nch = C / 8;
  if (nch <= 32) {
    if (nch & (nch - 1)) return false;
    *lpp = nch; *nck = 1;
    return true;
  }
  if (nch == 64 || nch == 128) { *lpp = 32; *nck = nch / 32; return true; }
  return false;
}

#define PG_PW_DISPATCH_LPP(lpp, MACRO)                                                     \
  switch (lpp) {                                                                           \
    case 1: MACRO(1, 1) break;                                                             \
    case 2: MACRO(2, 1) break;                                                             \
    case 4: MACRO(4, 1) break;                                                             \
    case 8: MACRO(8, 1) break;                                                             \
    case 16: MACRO(16, 1) break;                                                           \
    default: MACRO(32, 1) break;                                                           \
  }
// image-channel count as a template parameter: 1 (linear, mnist), 3 (rgb), 4 (rgb + label plane, 2)
#define PG_PW_DISPATCH_K(K, BODY)                                                          \
  if ((K) == 1) { constexpr int KT = 1; BODY }                                             \
  else if ((K) == 3) { constexpr int KT = 3; BODY }                                        \
  else { constexpr int KT = 4; BODY }
// the warp-run kernels index with 32 bits
static bool pw_fits32(long long P, int C, int K, long long HW, int w_sc, int w_sk) {
  return P * C < (1ll << 31) && P + 32 < (1ll << 31) && (long long)C * w_sc + (long long)K * w_sk < (1ll << 31)
         && HW < (1ll << 31);
}

extern "C" int pg_pw_expand(const float *img, const float *w, const float *bias, void *act, int N,
                            long long HW, int K, int C, int w_sc, int w_sk, float scale,
                            int dtype, void *stream) {
  PG_CHECK_ARG(img && w && act, "pg_pw_expand: null pointer");
  if (int rc = check_pw("pg_pw_expand", N, HW, K, C)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  int lpp = 0, nck = 0;
  if (pw_shape(C, &lpp, &nck) && pw_fits32((long long)N * HW, C, K, HW, w_sc, w_sk)) {
    const int P = (int)((long long)N * HW);
    const int grid = bw_grid((P + 31) / 32, 8 * 2);   // 8 warps x 2 runs per block pass
#define PG_PWE(L, NC) pg::launcher(pw_expand_kernel<T, L, NC, KT>, grid, 256, 0, s)(img, w, bias, (T *)act, P, (int)HW, K, w_sc, w_sk, scale);
    PG_DISPATCH_DTYPE(dtype, T, {
      PG_PW_DISPATCH_K(K, {
        if (nck == 1) { PG_PW_DISPATCH_LPP(lpp, PG_PWE) }
        else if (nck == 2) { PG_PWE(32, 2) }
        else { PG_PWE(32, 4) }
      })
    });
#undef PG_PWE
    PG_CHECK_LAUNCH("pg_pw_expand");
  }
  const long long total = (long long)N * HW * (C / 8);
  const int grid = bw_grid(total, 256);
  const size_t smem = (size_t)(K + 1) * C * sizeof(float);
  const bool small = total + (long long)grid * 256 < (1ll << 31);
  PG_DISPATCH_DTYPE(dtype, T, {
    if (small)
      pg::launcher(pw_expand_generic_kernel<T, unsigned>, grid, 256, smem, s)(
          img, w, bias, (T *)act, N, (unsigned)HW, K, C, w_sc, w_sk, scale);
    else
      pg::launcher(pw_expand_generic_kernel<T, long long>, grid, 256, smem, s)(
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
  const long long P = (long long)N * HW;
  cudaStream_t s = (cudaStream_t)stream;
  int lpp = 0, nck = 0;
  if (pw_shape(C, &lpp, &nck) && pw_fits32(P, C, K, HW, w_sc, w_sk)) {
    const int grid = bw_grid((P + 31) / 32, 8);
#define PG_PWR(L, NC) pg::launcher(pw_reduce_kernel<T, L, NC, KT>, grid, 256, 0, s)((const T *)act, w, bias, img, (int)P, (int)HW, K, w_sc, w_sk, scale);
    PG_DISPATCH_DTYPE(dtype, T, {
      PG_PW_DISPATCH_K(K, {
        if (nck == 1) { PG_PW_DISPATCH_LPP(lpp, PG_PWR) }
        else if (nck == 2) { PG_PWR(32, 2) }
        else { PG_PWR(32, 4) }
      })
    });
#undef PG_PWR
    PG_CHECK_LAUNCH("pg_pw_reduce");
  }
  const size_t smem = (size_t)K * C * sizeof(float);
#define PG_LAUNCH_PWR(TPP, CPL)                                                             \
  {                                                                                         \
    const int grid = bw_grid(P, 256 / TPP);                                                 \
    if (P + (long long)grid * 256 < (1ll << 31))                                            \
      pg::launcher(pw_reduce_generic_kernel<T, TPP, unsigned, CPL>, grid, 256, smem, s)(              \
          (const T *)act, w, bias, img, N, (unsigned)HW, K, C, w_sc, w_sk, scale);          \
    else                                                                                    \
      pg::launcher(pw_reduce_generic_kernel<T, TPP, long long, CPL>, grid, 256, smem, s)(             \
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

extern "C" int pg_colsum(const void *x, float *out, long long P, int C, int dtype, void *stream);

extern "C" int pg_pw_wgrad(const void *act, const float *img, float *dw, float *dbias, int N,
                           long long HW, int K, int C, int w_sc, int w_sk, float scale, int dtype,
                           void *stream) {
  PG_CHECK_ARG(img && dw && act, "pg_pw_wgrad: null pointer");
  if (int rc = check_pw("pg_pw_wgrad", N, HW, K, C)) return rc;
  const int nch = C / 8;
  const long long P = (long long)N * HW;
  cudaStream_t s = (cudaStream_t)stream;
  int lpp = 0, nck = 0;
  if (pw_shape(C, &lpp, &nck) && nck == 1 && pw_fits32(P, C, K, HW, w_sc, w_sk)) {
    // few enough blocks that the final atomics stay cheap, enough warps to cover the HBM latency
    const int grid = bw_grid((P + 31) / 32, 8 * 4, 4);
#define PG_PWW(L, NC) pg::launcher(pw_wgrad_kernel<T, L, KT>, grid, 256, 0, s)((const T *)act, img, dw, dbias, (int)P, (int)HW, K, w_sc, w_sk, scale);
    PG_DISPATCH_DTYPE(dtype, T, { PG_PW_DISPATCH_K(K, { PG_PW_DISPATCH_LPP(lpp, PG_PWW) }) });
#undef PG_PWW
    PG_CHECK_LAUNCH("pg_pw_wgrad");
  }
  const int rows = 256 / nch;
  PG_CHECK_ARG(rows >= 1, "pg_pw_wgrad: C too large");
  const int grid = bw_grid(P, rows * 16, 4);
  const size_t smem = (size_t)rows * K * C * sizeof(float);
  PG_DISPATCH_DTYPE(dtype, T, {
    if (P + (long long)grid * 256 < (1ll << 31))
      pg::launcher(pw_wgrad_generic_kernel<T, unsigned>, grid, 256, smem, s)(
          (const T *)act, img, dw, N, (unsigned)HW, K, C, w_sc, w_sk, scale);
    else
      pg::launcher(pw_wgrad_generic_kernel<T, long long>, grid, 256, smem, s)(
          (const T *)act, img, dw, N, HW, K, C, w_sc, w_sk, scale);
  });
  if (dbias) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("pg_pw_wgrad: CUDA launch failed: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    return pg_colsum(act, dbias, P, C, dtype, stream);
  }
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
  pg::launcher(img_chansum_kernel, grid, 256, 0, (cudaStream_t)stream)(img, out, N, HW, K);
  PG_CHECK_LAUNCH("pg_img_chansum");
}
