// Minibatch standard deviation (progan_modules.py:289-293) forward, backward and the
// batch-coupled second-order term needed by the WGAN-GP double backward.
//
//   mu_f = mean_n x_nf ; sigma_f = sqrt(mean_n (x_nf-mu_f)^2 + 1e-8) ; m = mean_f sigma_f
//   out  = cat([x, m * ones(N,1,4,4)], channel)           (channel-padded to Cp, zeros)
//
// The tensor is tiny ([N,4,4,C], 131k elements for N=64, C=128) and couples the whole
// batch, so one 1024-thread CTA owns it: thread t keeps the statistics of features
// f = t, t+1024, ... in registers, loads are coalesced over f, the scalar reductions
// are warp-shuffle block sums.  x is [N,F], F = 16*C, f = pos*C + c (NHWC).
#include "common.cuh"

namespace pg {

constexpr int kMbThreads = 1024;
constexpr int kMbMaxF = 8;  // F <= 8192  (C <= 512)

template <typename T>
struct MbStats {
  float mu[kMbMaxF], sigma[kMbMaxF];
  int nf;
  __device__ __forceinline__ void compute(const T *x, int N, int F) {
    nf = 0;
    for (int f = threadIdx.x; f < F; f += kMbThreads, ++nf) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += ldf(x + (long long)n * F + f);
      const float mean = s / (float)N;
      float v = 0.f;
      for (int n = 0; n < N; ++n) {
        const float d = ldf(x + (long long)n * F + f) - mean;
        v += d * d;
      }
      mu[nf] = mean;
      sigma[nf] = sqrtf(v / (float)N + 1e-8f);
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kMbThreads)
mbstd_fwd_kernel(const T *__restrict__ x, T *__restrict__ out, int N, int C, int Cp) {
  __shared__ float red[32];
  const int F = 16 * C;
  MbStats<T> st;
  st.compute(x, N, F);
  float local = 0.f;
  for (int j = 0; j < st.nf; ++j) local += st.sigma[j];
  const float m = block_sum(local, red) / (float)F;
  const long long total = (long long)N * 16 * Cp;
  for (long long i = threadIdx.x; i < total; i += kMbThreads) {
    const int c = (int)(i % Cp);
    const long long np = i / Cp;  // n*16 + pos
    float v;
    if (c < C) v = ldf(x + np * C + c);
    else v = (c == C) ? m : 0.f;
    stf(out + i, v);
  }
}

__device__ __forceinline__ float mb_delta_m(const float *red_in, float v, float *red) {
  return block_sum(v, red);
}

template <typename T>
__global__ void __launch_bounds__(kMbThreads)
mbstd_bwd_kernel(const T *__restrict__ dout, const T *__restrict__ x, T *__restrict__ dx,
                 int N, int C, int Cp) {
  __shared__ float red[32];
  const int F = 16 * C;
  float loc = 0.f;
  for (int i = threadIdx.x; i < N * 16; i += kMbThreads) loc += ldf(dout + (long long)i * Cp + C);
  const float dm = block_sum(loc, red);
  MbStats<T> st;
  st.compute(x, N, F);
  int j = 0;
  for (int f = threadIdx.x; f < F; f += kMbThreads, ++j) {
    const int pos = f / C, c = f - pos * C;
    const float k = dm / ((float)N * (float)F * st.sigma[j]);
    for (int n = 0; n < N; ++n) {
      const float xv = ldf(x + (long long)n * F + f);
      const float g = ldf(dout + ((long long)n * 16 + pos) * Cp + c);
      stf(dx + (long long)n * F + f, g + k * (xv - st.mu[j]));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kMbThreads)
mbstd_bwd_bwd_kernel(const T *__restrict__ t, const T *__restrict__ dout,
                     const T *__restrict__ x, T *__restrict__ cot_dout, T *__restrict__ cot_x,
                     int N, int C, int Cp) {
  __shared__ float red[32];
  const int F = 16 * C;
  float loc = 0.f;
  for (int i = threadIdx.x; i < N * 16; i += kMbThreads) loc += ldf(dout + (long long)i * Cp + C);
  const float dm = block_sum(loc, red);
  MbStats<T> st;
  st.compute(x, N, F);
  float tau_loc = 0.f;
  int j = 0;
  for (int f = threadIdx.x; f < F; f += kMbThreads, ++j) {
    float ts = 0.f, cs = 0.f;
    for (int n = 0; n < N; ++n) {
      const float tv = ldf(t + (long long)n * F + f);
      ts += tv;
      cs += tv * (ldf(x + (long long)n * F + f) - st.mu[j]);
    }
    const float tbar = ts / (float)N, cf = cs / (float)N;
    const float sg = st.sigma[j];
    tau_loc += cf / ((float)F * sg);
    const float k = dm / ((float)N * (float)F * sg);
    const float inv_var = 1.f / (sg * sg);
    for (int n = 0; n < N; ++n) {
      const float tv = ldf(t + (long long)n * F + f);
      const float xc = ldf(x + (long long)n * F + f) - st.mu[j];
      stf(cot_x + (long long)n * F + f, k * (tv - tbar - xc * cf * inv_var));
    }
  }
  const float tau = block_sum(tau_loc, red);
  const long long total = (long long)N * 16 * Cp;
  for (long long i = threadIdx.x; i < total; i += kMbThreads) {
    const int c = (int)(i % Cp);
    const long long np = i / Cp;
    float v;
    if (c < C) v = ldf(t + np * C + c);
    else v = (c == C) ? tau : 0.f;
    stf(cot_dout + i, v);
  }
}

}  // namespace pg

using namespace pg;

static int check_mb(const char *name, int N, int C, int Cp) {
  PG_CHECK_ARG(N > 0 && C > 0 && Cp > C, "%s: need N>0, Cp > C (N=%d C=%d Cp=%d)", name, N, C, Cp);
  PG_CHECK_ARG(16 * C <= kMbThreads * kMbMaxF, "%s: C=%d too large (max %d)", name, C,
               kMbThreads * kMbMaxF / 16);
  return PG_OK;
}

extern "C" int pg_mbstd_fwd(const void *x, void *out, int N, int C, int Cp, int dtype,
                            void *stream) {
  PG_CHECK_ARG(x && out, "pg_mbstd_fwd: null pointer");
  if (int rc = check_mb("pg_mbstd_fwd", N, C, Cp)) return rc;
  PG_DISPATCH_DTYPE(dtype, T, mbstd_fwd_kernel<T><<<1, kMbThreads, 0, (cudaStream_t)stream>>>(
                                  (const T *)x, (T *)out, N, C, Cp));
  PG_CHECK_LAUNCH("pg_mbstd_fwd");
}

extern "C" int pg_mbstd_bwd(const void *dout, const void *x, void *dx, int N, int C, int Cp,
                            int dtype, void *stream) {
  PG_CHECK_ARG(dout && x && dx, "pg_mbstd_bwd: null pointer");
  if (int rc = check_mb("pg_mbstd_bwd", N, C, Cp)) return rc;
  PG_DISPATCH_DTYPE(dtype, T, mbstd_bwd_kernel<T><<<1, kMbThreads, 0, (cudaStream_t)stream>>>(
                                  (const T *)dout, (const T *)x, (T *)dx, N, C, Cp));
  PG_CHECK_LAUNCH("pg_mbstd_bwd");
}

extern "C" int pg_mbstd_bwd_bwd(const void *t, const void *dout, const void *x, void *cot_dout,
                                void *cot_x, int N, int C, int Cp, int dtype, void *stream) {
  PG_CHECK_ARG(t && dout && x && cot_dout && cot_x, "pg_mbstd_bwd_bwd: null pointer");
  if (int rc = check_mb("pg_mbstd_bwd_bwd", N, C, Cp)) return rc;
  PG_DISPATCH_DTYPE(dtype, T,
                    mbstd_bwd_bwd_kernel<T><<<1, kMbThreads, 0, (cudaStream_t)stream>>>(
                        (const T *)t, (const T *)dout, (const T *)x, (T *)cot_dout, (T *)cot_x,
                        N, C, Cp));
  PG_CHECK_LAUNCH("pg_mbstd_bwd_bwd");
}
