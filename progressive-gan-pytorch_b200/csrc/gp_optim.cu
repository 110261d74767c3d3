// WGAN-GP scalar pieces (train.py:142-150), Adam (train.py:256-257) and the
// generator EMA (accumulate(), train.py:17-22).  All fp32, all HBM-bound:
// float4 accesses, warp-shuffle reductions, grids in multiples of the SM count.
#include "common.cuh"

namespace pg {

// x_hat[n,:] = eps[n]*real[n,:] + (1-eps[n])*fake[n,:]
__global__ void __launch_bounds__(256)
interp_xhat_kernel(const float *__restrict__ real, const float *__restrict__ fake,
                   const float *__restrict__ eps, float *__restrict__ out, int N, long long D) {
  pg::grid_dep_sync();
  const long long D4 = D >> 2;
  const long long total = (long long)N * D4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / D4, j = i - n * D4;
    const float e = eps[n];
    const float4 r = reinterpret_cast<const float4 *>(real + n * D)[j];
    const float4 f = reinterpret_cast<const float4 *>(fake + n * D)[j];
    float4 o;
    // same association and roundings as the reference expression eps*x + (1-eps)*G(z)
    // (train.py:143): two rounded products and one rounded sum, no FMA contraction,
    // so x_hat is bit-identical to torch's.
    const float e1 = __fsub_rn(1.f, e);
    o.x = __fadd_rn(__fmul_rn(e, r.x), __fmul_rn(e1, f.x));
    o.y = __fadd_rn(__fmul_rn(e, r.y), __fmul_rn(e1, f.y));
    o.z = __fadd_rn(__fmul_rn(e, r.z), __fmul_rn(e1, f.z));
    o.w = __fadd_rn(__fmul_rn(e, r.w), __fmul_rn(e1, f.w));
    reinterpret_cast<float4 *>(out + n * D)[j] = o;
  }
}

// one CTA per sample: norms[n] = ||g[n,:]||_2
__global__ void __launch_bounds__(1024)
gp_norm_kernel(const float *__restrict__ g, float *__restrict__ norms, long long D) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  const float *p = g + (long long)blockIdx.x * D;
  const long long D4 = D >> 2;
  float s = 0.f;
  for (long long j = threadIdx.x; j < D4; j += blockDim.x) {
    const float4 v = reinterpret_cast<const float4 *>(p)[j];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) norms[blockIdx.x] = sqrtf(s);
}

// Critic / generator losses of the loop body and their gradient seeds in one small launch
// (train.py:126-139,162-167): d = D outputs [n_real + n_fake] (real half first),
//   L = -(mean d_r - drift * mean d_r^2) + mean d_f      seed_n = dL/dd_n
//   metric[0] += mean d_r - drift * mean d_r^2 - mean d_f   (what the reference logs as disc loss)
// n_real = 0: generator form  L = -mean d,  metric[0] += L.
__global__ void __launch_bounds__(256)
wgan_loss_kernel(const float *__restrict__ d, float *__restrict__ seed, float *__restrict__ metric,
                 int n_real, int n_fake, float drift) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  float sr = 0.f, sr2 = 0.f, sf = 0.f;
  const float ir = n_real > 0 ? 1.f / (float)n_real : 0.f, ifk = 1.f / (float)n_fake;
  for (int n = threadIdx.x; n < n_real + n_fake; n += blockDim.x) {
    const float v = d[n];
    if (n < n_real) {
      sr += v;
      sr2 += v * v;
      seed[n] = (-1.f + 2.f * drift * v) * ir;
    } else {
      sf += v;
      seed[n] = n_real > 0 ? ifk : -ifk;
    }
  }
  sr = block_sum(sr, red);
  sr2 = block_sum(sr2, red);
  sf = block_sum(sf, red);
  if (threadIdx.x == 0 && metric != nullptr) {
    if (n_real > 0) *metric += sr * ir - drift * sr2 * ir - sf * ifk;
    else *metric += -sf * ifk;
  }
}

__global__ void gp_loss_kernel(const float *__restrict__ norms, float *__restrict__ gp, int N,
                               float lambda) {
  pg::grid_dep_sync();
  __shared__ float red[32];
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = norms[n] - 1.f;
    s += d * d;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) *gp = lambda * s / (float)N;
}

__global__ void __launch_bounds__(256)
gp_bwd_kernel(const float *__restrict__ g, const float *__restrict__ norms,
              const float *__restrict__ upstream, float *__restrict__ v, int N, long long D,
              float lambda) {
  pg::grid_dep_sync();
  const long long D4 = D >> 2;
  const long long total = (long long)N * D4;
  const float up = upstream ? *upstream : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / D4, j = i - n * D4;
    const float nm = norms[n];
    // torch's norm backward uses the subgradient 0 at ||g|| = 0 (no NaN into the gradient bucket)
    const float k = nm > 0.f ? up * 2.f * lambda / (float)N * (nm - 1.f) / nm : 0.f;
    float4 x = reinterpret_cast<const float4 *>(g + n * D)[j];
    x.x *= k; x.y *= k; x.z *= k; x.w *= k;
    reinterpret_cast<float4 *>(v + n * D)[j] = x;
  }
}

// Adam, torch.optim.Adam semantics (no weight decay, no amsgrad):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2
//   p -= lr / (1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, long long n, float lr, float b1, float b2, float eps,
            const float *__restrict__ step_dev, float grad_scale) {
  pg::grid_dep_sync();
  const float t = *step_dev;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float mi = gi;
    if (m) {
      mi = b1 * m[i] + (1.f - b1) * gi;
      m[i] = mi;
    }
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void __launch_bounds__(256)
ema_kernel(float *__restrict__ ema, const float *__restrict__ p, long long n, float decay) {
  pg::grid_dep_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    ema[i] = ema[i] * decay + (1.f - decay) * p[i];
}

}  // namespace pg

using namespace pg;

extern "C" int pg_interp_xhat(const float *real, const float *fake, const float *eps, float *out,
                              int N, long long D, void *stream) {
  PG_CHECK_ARG(real && fake && eps && out, "pg_interp_xhat: null pointer");
  PG_CHECK_ARG(N > 0 && D > 0 && D % 4 == 0, "pg_interp_xhat: need D %% 4 == 0");
  const int grid = bw_grid((long long)N * (D / 4), 256);
  pg::launcher(interp_xhat_kernel, grid, 256, 0, (cudaStream_t)stream)(real, fake, eps, out, N, D);
  PG_CHECK_LAUNCH("pg_interp_xhat");
}

extern "C" int pg_gp_fwd(const float *g, float *norms, float *gp, int N, long long D,
                         float lambda, void *stream) {
  PG_CHECK_ARG(g && norms && gp, "pg_gp_fwd: null pointer");
  PG_CHECK_ARG(N > 0 && D > 0 && D % 4 == 0, "pg_gp_fwd: need D %% 4 == 0");
  int threads = 1024;
  while (threads > 32 && (long long)threads * 4 > D) threads >>= 1;
  pg::launcher(gp_norm_kernel, N, threads, 0, (cudaStream_t)stream)(g, norms, D);
  pg::launcher(gp_loss_kernel, 1, 256, 0, (cudaStream_t)stream)(norms, gp, N, lambda);
  PG_CHECK_LAUNCH("pg_gp_fwd");
}

extern "C" int pg_gp_bwd(const float *g, const float *norms, const float *upstream, float *v,
                         int N, long long D, float lambda, void *stream) {
  PG_CHECK_ARG(g && norms && v, "pg_gp_bwd: null pointer");
  PG_CHECK_ARG(N > 0 && D > 0 && D % 4 == 0, "pg_gp_bwd: need D %% 4 == 0");
  const int grid = bw_grid((long long)N * (D / 4), 256);
  pg::launcher(gp_bwd_kernel, grid, 256, 0, (cudaStream_t)stream)(g, norms, upstream, v, N, D, lambda);
  PG_CHECK_LAUNCH("pg_gp_bwd");
}

extern "C" int pg_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr,
                            float beta1, float beta2, float eps, const float *step_dev,
                            float grad_scale, void *stream) {
  PG_CHECK_ARG(p && g && v && step_dev, "pg_adam_step: null pointer");
  PG_CHECK_ARG(m || beta1 == 0.f, "pg_adam_step: beta1 != 0 needs a first-moment buffer");
  PG_CHECK_ARG(n > 0, "pg_adam_step: n must be > 0");
  pg::launcher(adam_kernel, bw_grid(n, 256), 256, 0, (cudaStream_t)stream)(p, g, m, v, n, lr, beta1, beta2,
                                                                 eps, step_dev, grad_scale);
  PG_CHECK_LAUNCH("pg_adam_step");
}

extern "C" int pg_ema(float *ema, const float *p, long long n, float decay, void *stream) {
  PG_CHECK_ARG(ema && p && n > 0, "pg_ema: bad args");
  pg::launcher(ema_kernel, bw_grid(n, 256), 256, 0, (cudaStream_t)stream)(ema, p, n, decay);
  PG_CHECK_LAUNCH("pg_ema");
}

// ---- multi-tensor Adam over a flat bucket ------------------------------------------
// torch.optim.Adam keeps one step counter per parameter and skips parameters whose grad is
// None (inactive resolutions, SURVEY.md §7 "unused parameters per step").  The flat bucket
// is cut into chunks; chunk c belongs to segment seg[c] (a group of parameters that became
// active together) whose 1-based step count lives in steps_dev[seg] on the device.
namespace pg {
__global__ void __launch_bounds__(256)
adam_multi_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                  float *__restrict__ v, const int4 *__restrict__ chunks,
                  const float *__restrict__ steps_dev, float lr, float b1, float b2, float eps,
                  float grad_scale) {
  pg::grid_dep_sync();
  const int4 c = chunks[blockIdx.x];   // x = start, y = length, z = segment
  const float t = steps_dev[c.z];
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  // four independent elements per thread and iteration, every load issued before the arithmetic
  // (one element per iteration left a single 4-byte load in flight per thread: 0.29 of HBM)
  constexpr int UN = 4;
  for (int i0 = threadIdx.x; i0 < c.y; i0 += blockDim.x * UN) {
    float gi[UN], vi[UN], pi[UN], mi[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < c.y) {
        const long long j = (long long)c.x + i;
        gi[u] = g[j];
        vi[u] = v[j];
        pi[u] = p[j];
        mi[u] = m ? m[j] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < c.y) {
        const long long j = (long long)c.x + i;
        const float gs = gi[u] * grad_scale;
        float mm = gs;
        if (m) {
          mm = b1 * mi[u] + (1.f - b1) * gs;
          m[j] = mm;
        }
        const float vv = b2 * vi[u] + (1.f - b2) * gs * gs;
        v[j] = vv;
        p[j] = pi[u] - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
      }
    }
  }
}
}  // namespace pg

extern "C" int pg_adam_multi(float *p, const float *g, float *m, float *v, const void *chunks,
                             int nchunks, const float *steps_dev, float lr, float beta1,
                             float beta2, float eps, float grad_scale, void *stream) {
  PG_CHECK_ARG(p && g && v && chunks && steps_dev, "pg_adam_multi: null pointer");
  PG_CHECK_ARG(m || beta1 == 0.f, "pg_adam_multi: beta1 != 0 needs a first-moment buffer");
  PG_CHECK_ARG(nchunks > 0, "pg_adam_multi: nchunks must be > 0");
  pg::launcher(pg::adam_multi_kernel, nchunks, 256, 0, (cudaStream_t)stream)(
      p, g, m, v, (const int4 *)chunks, steps_dev, lr, beta1, beta2, eps, grad_scale);
  PG_CHECK_LAUNCH("pg_adam_multi");
}

extern "C" int pg_wgan_loss(const float *d, float *seed, float *metric, int n_real, int n_fake,
                            float drift, void *stream) {
  PG_CHECK_ARG(d && seed && n_real >= 0 && n_fake > 0, "pg_wgan_loss: bad arguments");
  pg::launcher(pg::wgan_loss_kernel, 1, 256, 0, (cudaStream_t)stream)(d, seed, metric, n_real, n_fake, drift);
  PG_CHECK_LAUNCH("pg_wgan_loss");
}
