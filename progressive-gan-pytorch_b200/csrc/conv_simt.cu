// Weight packing + generic kxk stride-1 convolution on CUDA cores (fp32 accumulate).
//
// This is the precision ("check mode") path and the path for shapes the tcgen05
// kernel does not take.  It replaces aten::convolution / convolution_backward as
// called through nn.Conv2d / nn.ConvTranspose2d at reference progan_modules.py:67,81.
#include "common.cuh"

namespace pg {

// ---------------------------------------------------------------- packing ----
template <typename T>
__global__ void pack_weight_kernel(const float *__restrict__ w, T *__restrict__ out,
                                   int d0, int d1, int taps, int swap_io, int flip,
                                   int layout, int ci_pad, int co_pad) {
  pg::grid_dep_sync();
  const int Cout = swap_io ? d1 : d0;
  const int Cin = swap_io ? d0 : d1;
  const long long total = (long long)co_pad * taps * ci_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co, ci, tap;
    if (layout == PG_WL_TAP_CI_CO) {
      co = (int)(i % co_pad);
      ci = (int)((i / co_pad) % ci_pad);
      tap = (int)(i / ((long long)co_pad * ci_pad));
    } else if (layout == PG_WL_CO_TAP_CI) {
      ci = (int)(i % ci_pad);
      tap = (int)((i / ci_pad) % taps);
      co = (int)(i / ((long long)ci_pad * taps));
    } else {  // PG_WL_TAP_CO_CI
      ci = (int)(i % ci_pad);
      co = (int)((i / ci_pad) % co_pad);
      tap = (int)(i / ((long long)ci_pad * co_pad));
    }
    float v = 0.f;
    if (ci < Cin && co < Cout) {
      const int st = flip ? (taps - 1 - tap) : tap;
      const long long src = swap_io ? ((long long)ci * d1 + co) * taps + st
                                    : ((long long)co * d1 + ci) * taps + st;
      v = w[src];
    }
    stf(out + i, v);
  }
}

// every (weight, operand layout) pair of a network in one launch: blockIdx.y = table entry
__global__ void __launch_bounds__(256)
pack_weight_multi_kernel(const PgPackEntry *__restrict__ table) {
  pg::grid_dep_sync();
  const PgPackEntry e = table[blockIdx.y];
  const int Cout = e.swap_io ? e.d1 : e.d0;
  const int Cin = e.swap_io ? e.d0 : e.d1;
  if (e.layout != PG_WL_TAP_CI_CO && e.taps <= 16) {
    // ci-fastest operand layouts (the tcgen05 kernels): one thread per (co, ci) pair reads the
    // taps of its pair as ONE contiguous run of the parameter and writes one element per tap,
    // coalesced over ci.  (One thread per output element read the parameter 4 bytes at a time,
    // `taps` floats apart: 0.11 of HBM.)
    const long long pairs = (long long)e.co_pad * e.ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pairs;
         i += (long long)gridDim.x * blockDim.x) {
      const int ci = (int)(i % e.ci_pad), co = (int)(i / e.ci_pad);
      float v[16];
      const bool ok = ci < Cin && co < Cout;
      const float *src = e.w + (e.swap_io ? ((long long)ci * e.d1 + co) : ((long long)co * e.d1 + ci)) * e.taps;
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (t < e.taps) v[t] = ok ? src[t] : 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (t < e.taps) {
          const int tap = e.flip ? (e.taps - 1 - t) : t;          // parameter tap t -> operand tap
          const long long o = e.layout == PG_WL_CO_TAP_CI ? ((long long)co * e.taps + tap) * e.ci_pad + ci
                                                          : ((long long)tap * e.co_pad + co) * e.ci_pad + ci;
          if (e.dtype == PG_BF16) stf(reinterpret_cast<__nv_bfloat16 *>(e.out) + o, v[t]);
          else stf(reinterpret_cast<float *>(e.out) + o, v[t]);
        }
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e.total;
       i += (long long)gridDim.x * blockDim.x) {
    int co, ci, tap;
    if (e.layout == PG_WL_TAP_CI_CO) {
      co = (int)(i % e.co_pad);
      ci = (int)((i / e.co_pad) % e.ci_pad);
      tap = (int)(i / ((long long)e.co_pad * e.ci_pad));
    } else if (e.layout == PG_WL_CO_TAP_CI) {
      ci = (int)(i % e.ci_pad);
      tap = (int)((i / e.ci_pad) % e.taps);
      co = (int)(i / ((long long)e.ci_pad * e.taps));
    } else {
      ci = (int)(i % e.ci_pad);
      co = (int)((i / e.ci_pad) % e.co_pad);
      tap = (int)(i / ((long long)e.ci_pad * e.co_pad));
    }
    float v = 0.f;
    if (ci < Cin && co < Cout) {
      const int st = e.flip ? (e.taps - 1 - tap) : tap;
      const long long src = e.swap_io ? ((long long)ci * e.d1 + co) * e.taps + st
                                      : ((long long)co * e.d1 + ci) * e.taps + st;
      v = e.w[src];
    }
    if (e.dtype == PG_BF16) stf(reinterpret_cast<__nv_bfloat16 *>(e.out) + i, v);
    else stf(reinterpret_cast<float *>(e.out) + i, v);
  }
}

// ---------------------------------------------------------------- forward ----
// one CTA per output pixel, one thread per output channel (loops if Cout > blockDim)
template <typename T>
__global__ void conv_fwd_simt_kernel(const T *__restrict__ x, const T *__restrict__ wp,
                                     const float *__restrict__ bias, T *__restrict__ y,
                                     float *__restrict__ r_out, int N, int H, int W,
                                     int Cin, int Ho, int Wo, int Cout, int k, int pad,
                                     float scale, int epi, float slope) {
  pg::grid_dep_sync();
  extern __shared__ float xs[];  // [k*k][Cin]
  __shared__ float red[32];
  const long long pix = blockIdx.x;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((long long)Wo * Ho));
  const int taps = k * k;
  for (int i = threadIdx.x; i < taps * Cin; i += blockDim.x) {
    const int tap = i / Cin, ci = i - tap * Cin;
    const int iy = oy + tap / k - pad, ix = ox + tap % k - pad;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = ldf(x + (((long long)n * H + iy) * W + ix) * Cin + ci);
    xs[i] = v;
  }
  __syncthreads();
  // each thread owns channels co = tid, tid+blockDim, ... (at most 4 supported)
  float a[4];
  float ss = 0.f;
  int nown = 0;
  for (int co = threadIdx.x; co < Cout; co += blockDim.x, ++nown) {
    float acc = 0.f;
    const T *wcol = wp + co;
    for (int i = 0; i < taps * Cin; ++i) acc = fmaf(xs[i], ldf(wcol + (long long)i * Cout), acc);
    float v = acc * scale + (bias ? bias[co] : 0.f);
    a[nown] = v;
    ss += v * v;
  }
  float r = 1.f;
  if (epi == PG_EPI_PN_LRELU) {
    ss = block_sum(ss, red);
    r = rsqrtf(ss / (float)Cout + 1e-8f);
    if (threadIdx.x == 0) r_out[pix] = r;
  }
  nown = 0;
  for (int co = threadIdx.x; co < Cout; co += blockDim.x, ++nown) {
    float v = a[nown] * r;
    if (epi != PG_EPI_LINEAR) v = v > 0.f ? v : v * slope;
    stf(y + pix * Cout + co, v);
  }
}

// ------------------------------------------------------------ weight grad ----
// grid: (ceil(Cin/32), ceil(Cout/8), taps*splits); block (32, 8)
template <typename T>
__global__ void conv_wgrad_simt_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                       float *__restrict__ dw, int N, int H, int W, int Cin,
                                       int Ho, int Wo, int Cout, int k, int pad, float scale,
                                       int swap_io, int flip, int splits, int d1) {
  pg::grid_dep_sync();
  const int ci = blockIdx.x * 32 + threadIdx.x;
  const int co = blockIdx.y * 8 + threadIdx.y;
  const int taps = k * k;
  const int tap = blockIdx.z % taps;
  const int split = blockIdx.z / taps;
  const long long P = (long long)N * Ho * Wo;
  const long long per = (P + splits - 1) / splits;
  const long long p0 = split * per, p1 = (p0 + per < P) ? p0 + per : P;
  const int u = tap / k, v = tap % k;
  float acc = 0.f;
  const bool active = (ci < Cin) && (co < Cout);
  for (long long p = p0; p < p1; ++p) {
    const int ox = (int)(p % Wo);
    const int oy = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((long long)Wo * Ho));
    const int iy = oy + u - pad, ix = ox + v - pad;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    if (active) {
      const float g = ldf(dy + p * Cout + co);
      const float xv = ldf(x + (((long long)n * H + iy) * W + ix) * Cin + ci);
      acc = fmaf(g, xv, acc);
    }
  }
  if (active) {
    const int st = flip ? (taps - 1 - tap) : tap;
    const long long dst = swap_io ? ((long long)ci * d1 + co) * taps + st
                                  : ((long long)co * d1 + ci) * taps + st;
    atomicAdd(dw + dst, acc * scale);
  }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_pack_conv_weight(const float *w, void *out, int d0, int d1, int kh, int kw,
                                   int swap_io, int flip, int out_layout, int ci_pad, int co_pad,
                                   int out_dtype, void *stream) {
  const int Cin = swap_io ? d0 : d1;
  PG_CHECK_ARG(w && out, "pg_pack_conv_weight: null pointer");
  PG_CHECK_ARG(d0 > 0 && d1 > 0 && kh > 0 && kw > 0, "pg_pack_conv_weight: bad dims");
  PG_CHECK_ARG(ci_pad >= Cin, "pg_pack_conv_weight: ci_pad %d < Cin %d", ci_pad, Cin);
  PG_CHECK_ARG(out_layout >= PG_WL_TAP_CI_CO && out_layout <= PG_WL_TAP_CO_CI,
               "pg_pack_conv_weight: bad layout %d", out_layout);
  const int Cout = swap_io ? d1 : d0;
  PG_CHECK_ARG(co_pad >= Cout, "pg_pack_conv_weight: co_pad %d < Cout %d", co_pad, Cout);
  const long long total = (long long)co_pad * kh * kw * ci_pad;
  const int grid = bw_grid(total, 256);
  PG_DISPATCH_DTYPE(out_dtype, T,
                    pg::launcher(pack_weight_kernel<T>, grid, 256, 0, (cudaStream_t)stream)(
                        w, (T *)out, d0, d1, kh * kw, swap_io, flip, out_layout, ci_pad, co_pad));
  PG_CHECK_LAUNCH("pg_pack_conv_weight");
}

extern "C" int pg_pack_conv_weight_multi(const PgPackEntry *table, int n, void *stream) {
  PG_CHECK_ARG(table && n > 0 && n <= 65535, "pg_pack_conv_weight_multi: bad table");
  dim3 grid(64, (unsigned)n);
  pg::launcher(pack_weight_multi_kernel, grid, 256, 0, (cudaStream_t)stream)(table);
  PG_CHECK_LAUNCH("pg_pack_conv_weight_multi");
}

extern "C" int pg_conv_fwd_simt(const void *x, const void *wp, const float *bias, void *y,
                                float *r_out, int N, int H, int W, int Cin, int Cout, int k,
                                int pad, float scale, int epi, float slope, int dtype,
                                void *stream) {
  PG_CHECK_ARG(x && wp && y, "pg_conv_fwd_simt: null pointer");
  PG_CHECK_ARG(epi != PG_EPI_PN_LRELU || r_out, "pg_conv_fwd_simt: PN epilogue needs r_out");
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  PG_CHECK_ARG(N > 0 && Ho > 0 && Wo > 0 && Cin > 0 && Cout > 0, "pg_conv_fwd_simt: bad dims");
  int threads = ((Cout + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  PG_CHECK_ARG(Cout <= 4 * threads, "pg_conv_fwd_simt: Cout %d > 1024 unsupported", Cout);
  const size_t smem = (size_t)k * k * Cin * sizeof(float);
  PG_CHECK_ARG(smem <= 200 * 1024, "pg_conv_fwd_simt: k*k*Cin too large for shared memory");
  const long long P = (long long)N * Ho * Wo;
  PG_CHECK_ARG(P < (1ll << 31), "pg_conv_fwd_simt: too many pixels");
  PG_DISPATCH_DTYPE(dtype, T, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
    pg::launcher(conv_fwd_simt_kernel<T>, (unsigned)P, threads, smem, (cudaStream_t)stream)(
        (const T *)x, (const T *)wp, bias, (T *)y, r_out, N, H, W, Cin, Ho, Wo, Cout, k, pad,
        scale, epi, slope);
  });
  PG_CHECK_LAUNCH("pg_conv_fwd_simt");
}

extern "C" int pg_conv_wgrad_simt(const void *x, const void *dy, float *dw, int N, int H, int W,
                                  int Cin, int Cout, int k, int pad, float scale, int swap_io,
                                  int flip, int dtype, void *stream) {
  PG_CHECK_ARG(x && dy && dw, "pg_conv_wgrad_simt: null pointer");
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  PG_CHECK_ARG(N > 0 && Ho > 0 && Wo > 0 && Cin > 0 && Cout > 0, "pg_conv_wgrad_simt: bad dims");
  const long long P = (long long)N * Ho * Wo;
  int splits = (int)((P + 2047) / 2048);
  if (splits < 1) splits = 1;
  if (splits > 512) splits = 512;
  const int taps = k * k;
  PG_CHECK_ARG((long long)taps * splits <= 65535, "pg_conv_wgrad_simt: grid.z overflow");
  dim3 grid((Cin + 31) / 32, (Cout + 7) / 8, taps * splits), block(32, 8);
  const int d1 = swap_io ? Cout : Cin;
  PG_DISPATCH_DTYPE(dtype, T,
                    pg::launcher(conv_wgrad_simt_kernel<T>, grid, block, 0, (cudaStream_t)stream)(
                        (const T *)x, (const T *)dy, dw, N, H, W, Cin, Ho, Wo, Cout, k, pad,
                        scale, swap_io, flip, splits, d1));
  PG_CHECK_LAUNCH("pg_conv_wgrad_simt");
}
