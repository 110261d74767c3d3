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

// MAXI = chunks of 8 channels per lane (C <= 8*TPP*MAXI)
template <typename T, int TPP, int MAXI, bool SECOND>
__global__ void __launch_bounds__(256)
pn_lrelu_grad_kernel(const T *__restrict__ t_in, const T *__restrict__ dy,
                     const T *__restrict__ y, const float *__restrict__ rr,
                     T *__restrict__ out0, T *__restrict__ out1, long long P, int C,
                     float slope, int use_pn) {
  // first order : out0 = da                 (t_in unused)
  // second order: out0 = cot_dy, out1 = cot_a
  const int sub = threadIdx.x % TPP;
  const long long ppb = blockDim.x / TPP;
  const int nch = C >> 3;
  const float invC = 1.f / (float)C;
  const float inv_slope = 1.f / slope;
  for (long long pix = blockIdx.x * ppb + threadIdx.x / TPP;
       pix < ((P + ppb - 1) / ppb) * ppb; pix += (long long)gridDim.x * ppb) {
    const bool live = pix < P;
    F8 pv[MAXI], uv[MAXI], tv[MAXI];
    float mk[MAXI][8];
    float s_pu = 0.f, s_pt = 0.f, s_tu = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int ch = sub + i * TPP;
      if (live && ch < nch) {
        const long long off = pix * C + (long long)ch * 8;
        F8 yv = ld8(y + off);
        F8 dv = ld8(dy + off);
        if (SECOND) tv[i] = ld8(t_in + off);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool pos = yv.v[e] > 0.f;
          const float m = pos ? 1.f : slope;
          const float p = pos ? yv.v[e] : yv.v[e] * inv_slope;
          mk[i][e] = m;
          pv[i].v[e] = p;
          uv[i].v[e] = m * dv.v[e];
          s_pu += p * uv[i].v[e];
          if (SECOND) {
            s_pt += p * tv[i].v[e];
            s_tu += tv[i].v[e] * uv[i].v[e];
          }
        }
      }
    }
    float r = 1.f;
    if (use_pn) {
      s_pu = subwarp_sum<TPP>(s_pu);
      if (SECOND) {
        s_pt = subwarp_sum<TPP>(s_pt);
        s_tu = subwarp_sum<TPP>(s_tu);
      }
      if (live) r = rr[pix];
    }
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int ch = sub + i * TPP;
      if (live && ch < nch) {
        const long long off = pix * C + (long long)ch * 8;
        F8 o0, o1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float p = pv[i].v[e], u = uv[i].v[e];
          if (!SECOND) {
            o0.v[e] = use_pn ? r * (u - p * s_pu * invC) : u;
          } else {
            const float t = tv[i].v[e];
            if (use_pn) {
              o0.v[e] = mk[i][e] * r * (t - p * s_pt * invC);
              o1.v[e] = r * r * invC *
                        (3.f * invC * s_pt * s_pu * p - s_tu * p - s_pu * t - s_pt * u);
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

template <typename T, bool SECOND>
static int launch_pn_grad(const T *t, const T *dy, const T *y, const float *r, T *o0, T *o1,
                          long long P, int C, float slope, int use_pn, cudaStream_t s) {
  const int nch = C / 8;
#define PG_LAUNCH_PN(TPP, MAXI)                                                           \
  {                                                                                       \
    const long long ppb = 256 / TPP;                                                      \
    const int grid = bw_grid(P, (int)ppb);                                                \
    pn_lrelu_grad_kernel<T, TPP, MAXI, SECOND><<<grid, 256, 0, s>>>(t, dy, y, r, o0, o1, \
                                                                     P, C, slope, use_pn); \
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
  extern __shared__ float sm[];  // [rows][C]
  const int nch = C >> 3;
  const int rows = blockDim.x / nch;
  const int cj = threadIdx.x % nch, rj = threadIdx.x / nch;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rj < rows) {
    for (long long pix = (long long)blockIdx.x * rows + rj; pix < P;
         pix += (long long)gridDim.x * rows) {
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

extern "C" int pg_pn_lrelu_bwd(const void *dy, const void *y, const float *r, void *da,
                               long long P, int C, float slope, int use_pn, int dtype,
                               void *stream) {
  PG_CHECK_ARG(dy && y && da && (r || !use_pn), "pg_pn_lrelu_bwd: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0, "pg_pn_lrelu_bwd: need C %% 8 == 0 (C=%d)", C);
  PG_CHECK_ARG(slope > 0.f, "pg_pn_lrelu_bwd: slope must be > 0");
  PG_DISPATCH_DTYPE(dtype, T, {
    int rc = launch_pn_grad<T, false>(nullptr, (const T *)dy, (const T *)y, r, (T *)da, nullptr,
                                      P, C, slope, use_pn, (cudaStream_t)stream);
    if (rc) return rc;
  });
  PG_CHECK_LAUNCH("pg_pn_lrelu_bwd");
}

extern "C" int pg_pn_lrelu_bwd_bwd(const void *t, const void *dy, const void *y, const float *r,
                                   void *cot_dy, void *cot_a, long long P, int C, float slope,
                                   int use_pn, int dtype, void *stream) {
  PG_CHECK_ARG(t && dy && y && cot_dy && cot_a && (r || !use_pn),
               "pg_pn_lrelu_bwd_bwd: null pointer");
  PG_CHECK_ARG(P > 0 && C > 0 && C % 8 == 0, "pg_pn_lrelu_bwd_bwd: need C %% 8 == 0 (C=%d)", C);
  PG_CHECK_ARG(slope > 0.f, "pg_pn_lrelu_bwd_bwd: slope must be > 0");
  PG_DISPATCH_DTYPE(dtype, T, {
    int rc = launch_pn_grad<T, true>((const T *)t, (const T *)dy, (const T *)y, r, (T *)cot_dy,
                                     (T *)cot_a, P, C, slope, use_pn, (cudaStream_t)stream);
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
  PG_DISPATCH_DTYPE(dtype, T, colsum_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>(
                                  (const T *)x, out, P, C));
  PG_CHECK_LAUNCH("pg_colsum");
}
