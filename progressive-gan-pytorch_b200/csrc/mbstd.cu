// Minibatch standard deviation (progan_modules.py:289-293) forward, backward and the
// batch-coupled second-order term needed by the WGAN-GP double backward.
//
//   mu_f = mean_n x_nf ; sigma_f = sqrt(mean_n (x_nf-mu_f)^2 + 1e-8) ; m = mean_f sigma_f
//   out  = cat([x, m * ones(N,1,4,4)], channel)           (channel-padded to Cp, zeros)
//
// x is [N,F], F = 16*C, f = pos*C + c (NHWC).  Two small launches per op: a statistics
// kernel (one thread per feature, coalesced over f, 128-thread CTAs so the F = 2048 features
// spread over 16 SMs; per-CTA partial sums of the scalar reductions) and a fully parallel
// elementwise kernel that finishes the scalar (sums the <= 64 partials itself — no atomics,
// no memset, deterministic).  The statistics (mu, sigma) travel in a small fp32 workspace.
#include "common.cuh"

namespace pg {

constexpr int kMbT = 128;      // threads per statistics CTA
constexpr int kMbMaxParts = 64;

// stats layout: [0,F) mu, [F,2F) sigma, [2F,3F) aux0, [3F,4F) aux1, [4F, 4F+64) partials
template <typename T>
__global__ void __launch_bounds__(kMbT)
mbstd_stats_kernel(const T *__restrict__ x, const T *__restrict__ t, float *__restrict__ stats, int N,
                   int F) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  const int f = blockIdx.x * kMbT + threadIdx.x;
  float part = 0.f;
  if (f < F) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += ldf(x + (long long)n * F + f);
    const float mean = s / (float)N;
    float v = 0.f, ts = 0.f, cs = 0.f;
    for (int n = 0; n < N; ++n) {
      const float d = ldf(x + (long long)n * F + f) - mean;
      v += d * d;
      if (t) {
        const float tv = ldf(t + (long long)n * F + f);
        ts += tv;
        cs += tv * d;
      }
    }
    const float sg = sqrtf(v / (float)N + 1e-8f);
    stats[f] = mean;
    stats[F + f] = sg;
    if (t) {
      const float cf = cs / (float)N;
      stats[2 * F + f] = ts / (float)N;   // tbar_f
      stats[3 * F + f] = cf;              // c_f
      part = cf / ((float)F * sg);        // contribution to tau
    } else {
      part = sg;                          // contribution to m*F
    }
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) stats[4 * F + blockIdx.x] = part;
}

__device__ __forceinline__ float mb_sum_parts(const float *stats, int F, int nparts) {
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += stats[4 * F + i];
  return s;
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_fwd_write_kernel(const T *__restrict__ x, const float *__restrict__ stats, T *__restrict__ out,
                       int N, int C, int Cp, int nparts) {
  pg::grid_dep_sync();
  const int F = 16 * C;
  const float m = mb_sum_parts(stats, F, nparts) / (float)F;
  const long long total = (long long)N * 16 * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long np = i / Cp;
    float v;
    if (c < C) v = ldf(x + np * C + c);
    else v = (c == C) ? m : 0.f;
    stf(out + i, v);
  }
}

// delta_m = sum over the N*16 entries of the statistic channel of dout (every CTA redoes it)
template <typename T>
__device__ __forceinline__ float mb_delta_m(const T *dout, int N, int C, int Cp, float *red) {
  float loc = 0.f;
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) loc += ldf(dout + (long long)i * Cp + C);
  return block_sum(loc, red);
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_bwd_kernel(const T *__restrict__ dout, const T *__restrict__ x, const float *__restrict__ stats,
                 T *__restrict__ dx, int N, int C, int Cp) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  const int F = 16 * C;
  const float dm = mb_delta_m(dout, N, C, Cp, red);
  const float k0 = dm / ((float)N * (float)F);
  const long long total = (long long)N * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long n = i / F;
    const int pos = f / C, c = f - pos * C;
    const float g = ldf(dout + (n * 16 + pos) * Cp + c);
    stf(dx + i, g + k0 / stats[F + f] * (ldf(x + i) - stats[f]));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_bwd_bwd_kernel(const T *__restrict__ t, const T *__restrict__ dout, const T *__restrict__ x,
                     const float *__restrict__ stats, T *__restrict__ cot_dout,
                     T *__restrict__ cot_x, int N, int C, int Cp, int nparts) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  const int F = 16 * C;
  const float dm = mb_delta_m(dout, N, C, Cp, red);
  const float tau = mb_sum_parts(stats, F, nparts);
  const float k0 = dm / ((float)N * (float)F);
  const long long total = (long long)N * 16 * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long np = i / Cp;       // n*16 + pos
    float v;
    if (c < C) {
      const int f = (int)(np % 16) * C + c;
      const long long j = np * C + c;
      const float tv = ldf(t + j);
      const float sg = stats[F + f];
      const float xc = ldf(x + j) - stats[f];
      stf(cot_x + j, k0 / sg * (tv - stats[2 * F + f] - xc * stats[3 * F + f] / (sg * sg)));
      v = tv;
    } else {
      v = (c == C) ? tau : 0.f;
    }
    stf(cot_dout + i, v);
  }
}

}  // namespace pg

using namespace pg;

static int check_mb(const char *name, int N, int C, int Cp) {
  PG_CHECK_ARG(N > 0 && C > 0 && Cp > C, "%s: need N>0, Cp > C (N=%d C=%d Cp=%d)", name, N, C, Cp);
  PG_CHECK_ARG((16 * C + kMbT - 1) / kMbT <= kMbMaxParts, "%s: C=%d too large (max %d)", name, C,
               kMbT * kMbMaxParts / 16);
  return PG_OK;
}

extern "C" int pg_mbstd_fwd(const void *x, void *out, float *stats, int N, int C, int Cp, int dtype,
                            void *stream) {
  PG_CHECK_ARG(x && out && stats, "pg_mbstd_fwd: null pointer");
  if (int rc = check_mb("pg_mbstd_fwd", N, C, Cp)) return rc;
  const int F = 16 * C, nparts = (F + kMbT - 1) / kMbT;
  cudaStream_t s = (cudaStream_t)stream;
  PG_DISPATCH_DTYPE(dtype, T, {
    pg::launcher(mbstd_stats_kernel<T>, nparts, kMbT, 0, s)((const T *)x, (const T *)nullptr, stats, N, F);
    pg::launcher(mbstd_fwd_write_kernel<T>, bw_grid((long long)N * 16 * Cp, 256, 2), 256, 0, s)(
        (const T *)x, stats, (T *)out, N, C, Cp, nparts);
  });
  PG_CHECK_LAUNCH("pg_mbstd_fwd");
}

extern "C" int pg_mbstd_bwd(const void *dout, const void *x, const float *stats, void *dx, int N,
                            int C, int Cp, int dtype, void *stream) {
  PG_CHECK_ARG(dout && x && dx && stats, "pg_mbstd_bwd: null pointer");
  if (int rc = check_mb("pg_mbstd_bwd", N, C, Cp)) return rc;
  PG_DISPATCH_DTYPE(dtype, T,
                    pg::launcher(mbstd_bwd_kernel<T>, bw_grid((long long)N * 16 * C, 256, 2), 256, 0,
                                          (cudaStream_t)stream)((const T *)dout, (const T *)x, stats,
                                                                  (T *)dx, N, C, Cp));
  PG_CHECK_LAUNCH("pg_mbstd_bwd");
}

extern "C" int pg_mbstd_bwd_bwd(const void *t, const void *dout, const void *x, float *stats,
                                void *cot_dout, void *cot_x, int N, int C, int Cp, int dtype,
                                void *stream) {
  PG_CHECK_ARG(t && dout && x && cot_dout && cot_x && stats, "pg_mbstd_bwd_bwd: null pointer");
  if (int rc = check_mb("pg_mbstd_bwd_bwd", N, C, Cp)) return rc;
  const int F = 16 * C, nparts = (F + kMbT - 1) / kMbT;
  cudaStream_t s = (cudaStream_t)stream;
  PG_DISPATCH_DTYPE(dtype, T, {
    pg::launcher(mbstd_stats_kernel<T>, nparts, kMbT, 0, s)((const T *)x, (const T *)t, stats, N, F);
    pg::launcher(mbstd_bwd_bwd_kernel<T>, bw_grid((long long)N * 16 * Cp, 256, 2), 256, 0, s)(
        (const T *)t, (const T *)dout, (const T *)x, stats, (T *)cot_dout, (T *)cot_x, N, C, Cp, nparts);
  });
  PG_CHECK_LAUNCH("pg_mbstd_bwd_bwd");
}
