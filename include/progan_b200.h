/*
 * progan_b200.h — C-ABI of the B200-native progressive-GAN training-step kernels.
 *
 * This is the drop-in boundary.  Every entry point replaces one
 * ATen/cuDNN library call (or a chain of unfused elementwise calls) that the
 * reference makes from progan_modules.py / train.py.  The reference's own
 * operator-plugin convention is ada/torch_utils/ops/bias_act.cpp:32-97 (pybind,
 * torch::Tensor in the signature); here the same role is played by plain
 * pointers + sizes so any host language can bind it.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller; nothing is retained,
 *     nothing is allocated (workspaces are passed in);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, negative = error, message via pg_last_error()
 *     (thread-local); there is NO CPU fallback anywhere in this library;
 *   - `dtype` of activation tensors: PG_F32 (check mode) or PG_BF16 (product
 *     mode).  Parameters, gradients of parameters, per-pixel statistics and
 *     all reductions are fp32.
 *   - feature maps are NHWC ("act" tensors); images at the module boundary are
 *     NCHW fp32 exactly as the reference train scripts pass them.
 */
#ifndef PROGAN_B200_H
#define PROGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PG_F32 0
#define PG_BF16 1

#define PG_OK 0
#define PG_ERR_INVALID (-1)
#define PG_ERR_CUDA (-2)
#define PG_ERR_UNSUPPORTED (-3)

/* epilogue modes of the conv kernels */
#define PG_EPI_LINEAR 0   /* y = s*conv + bias                                    */
#define PG_EPI_PN_LRELU 1 /* a = s*conv + bias; r = rsqrt(mean_c a^2 + 1e-8);     */
                          /* y = lrelu(a*r); r stored per pixel                   */
#define PG_EPI_LRELU 2    /* y = lrelu(s*conv + bias)   (pixel_norm=False blocks) */

/* packed-weight layouts */
#define PG_WL_TAP_CI_CO 0 /* [tap][ci][co]      — SIMT kernels                    */
#define PG_WL_CO_TAP_CI 1 /* [co][tap][ci]      — K-major rows for tcgen05        */
#define PG_WL_TAP_CO_CI 2 /* [tap][co][ci]      — GEMM form of a full conv on 1x1 */

const char *pg_last_error(void);
int pg_abi_version(void);
/* sm count / compute capability of the current device (negative on error). */
int pg_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- EqualLR weight handling: progan_modules.py:22-27,43-45 -------------
 * The equalized-LR scale sqrt(2/fan_in) is NOT applied here; it is folded
 * into every kernel's epilogue as `scale`.  Packing only re-lays
 * weight_orig (fp32, [d0][d1][kh][kw]) into the operand layout of the
 * "logical" cross-correlation  y[co] = sum Wl[co][ci][tap] x[ci]:
 *   swap_io=0: co=d0, ci=d1 (nn.Conv2d OIHW)   swap_io=1: co=d1, ci=d0
 *   flip=1   : tap -> kh*kw-1-tap (data-gradient / ConvTranspose forms)     */
/* ci_pad / co_pad >= logical Cin / Cout: extra rows/columns are zero-filled (activation
 * tensors whose channel count is padded to a multiple of 32 for the tensor-core path). */
int pg_pack_conv_weight(const float *w, void *out, int d0, int d1, int kh, int kw,
                        int swap_io, int flip, int out_layout, int ci_pad, int co_pad,
                        int out_dtype, void *stream);
/* The same re-layout for n weights in ONE launch (after an optimiser step every operand copy
 * of a network is refreshed at once).  `table` is a DEVICE array of n entries. */
typedef struct PgPackEntry {
  const float *w;
  void *out;
  long long total; /* co_pad * taps * ci_pad */
  int d0, d1, taps, swap_io, flip, layout, ci_pad, co_pad, dtype, reserved;
} PgPackEntry;
int pg_pack_conv_weight_multi(const PgPackEntry *table, int n, void *stream);

/* ---- generic kxk stride-1 conv, SIMT fp32-accumulate ----------------------
 * replaces aten::convolution for nn.Conv2d / nn.ConvTranspose2d at
 * progan_modules.py:67,81 in check mode and for shapes the tcgen05 kernel
 * does not take.  x:[N,H,W,Cin] act, wp: PG_WL_TAP_CI_CO, y:[N,Ho,Wo,Cout],
 * Ho = H + 2*pad - k + 1.  r_out:[N*Ho*Wo] fp32 (only PG_EPI_PN_LRELU).     */
int pg_conv_fwd_simt(const void *x, const void *wp, const float *bias, void *y,
                     float *r_out, int N, int H, int W, int Cin, int Cout, int k,
                     int pad, float scale, int epi, float slope, int dtype,
                     void *stream);
/* weight gradient of the same conv (aten::convolution_backward, weight part):
 * dw (fp32, [d0][d1][k][k] in the layout given by swap_io/flip, must be
 * zero-initialised by the caller) += scale * sum dy[n,o,co] x[n,o+tap-pad,ci] */
int pg_conv_wgrad_simt(const void *x, const void *dy, float *dw, int N, int H,
                       int W, int Cin, int Cout, int k, int pad, float scale,
                       int swap_io, int flip, int dtype, void *stream);

/* ---- implicit-GEMM conv on tcgen05 (bf16 in, fp32 accumulate) -------------
 * x:[N,H,W,Cin] bf16 (Cin % 32 == 0), wp: K-major bf16 [Cout_total][taps*Cin],
 * y:[N,H,W,Cout_total] bf16, tiled over N in tiles of Cout_tile (multiple of 32,
 * <= 256).  taps == 9: 3x3 pad 1.  taps == 1: 1x1 conv / plain GEMM — the 4x4
 * valid conv of D's last block (x viewed as [N,1,1,16*C]) and the 4x4
 * ConvTranspose of G's input layer (y viewed as [N,1,1,16*C]) run through this
 * form.  Same epilogues as the SIMT kernel; PixelNorm is taken over one N tile;
 * bias[c % bias_mod]; r_out:[N*H*W*n_tiles].  Layers wider than 256 channels
 * (512: the Correct* defaults) run with several N tiles of 256, bias_mod a
 * multiple of Cout_tile (tile t adds bias[(t*Cout_tile + c) % bias_mod]) and a
 * LINEAR/LRELU epilogue; their PixelNorm is pg_pn_lrelu_fwd.                   */
int pg_conv_tc(const void *x, const void *wp, const float *bias, void *y,
               float *r_out, int N, int H, int W, int Cin, int Cout_total,
               int Cout_tile, int taps, int bias_mod, float scale, int epi,
               float slope, void *y_pool, void *stream);
/* y_pool (nullable, [N,H/2,W/2,Cout]): the 2x2 average pool of y (the x0.5 bilinear downsample
 * after every discriminator block, progan_modules.py:299) written by the same epilogue; 3x3
 * form with H %% 16 == 0, W %% 8 == 0, Cout in {32,64,128} only. */
/* Data-gradient 3x3 conv with the PixelNorm+LeakyReLU backward of the layer in front fused into
 * the epilogue (replaces aten::convolution_backward(input) followed by the autograd chain of
 * PixelNorm/LeakyReLU, progan_modules.py:54-60,138):
 *   dh = scale * conv3x3(x; wp)            x:[N,H,W,Cin] = da of this layer, wp packed adjoint
 *   da_prev = Jpn(a_prev)^T (m * dh)       (y_prev, r_prev) = stored activation of the layer
 *                                          in front, [N,H,W,Cout] bf16 / [N,H,W] fp32
 *   colsum[c] += sum_pix da_prev[pix,c]    (bias gradient of the layer in front) if non-null
 * Returns PG_ERR_UNSUPPORTED for shapes the fused kernel does not take (H %% 16, W %% 8,
 * Cout in {32,64,128}, Cin %% 32). */
int pg_conv_tc_actbwd(const void *x, const void *wp, void *da, int N, int H, int W, int Cin,
                      int Cout, float scale, const void *y_prev, const float *r_prev,
                      float slope, int use_pn, float *colsum, void *stream);
/* weight gradient on tcgen05: dw fp32 (logical dims Cin_log/Cout_log, parameter
 * layout by swap_io/flip) is overwritten; workspace is taps*Cin*Cout floats.
 * Cout a multiple of 32 up to 256, or of 256 up to 1024 (dy tiles of 256 channels).
 * flat == 0: 3x3 pad 1 (taps == 9).  flat == 1: x is [N,1,1,taps*Cin] and tap t
 * addresses channel block t (weight gradient of the GEMM forms above).
 * accumulate == 1: dw += result (gradient accumulation straight into the flat bucket).
 * accumulate == 2: deferred — the partial sums are ADDED to `workspace` (which the caller keeps
 * zero-initialised and persistent) and dw is not touched; pg_wgrad_unpack_multi later folds
 * every pending workspace into its gradient in one launch. */
int pg_conv_wgrad_tc(const void *x, const void *dy, float *dw, float *workspace, int N,
                     int H, int W, int Cin, int Cout, int Cin_log, int Cout_log,
                     int taps, int flat, float scale, int swap_io, int flip,
                     int accumulate, void *stream);
/* Deferred weight-gradient epilogue for n parameters in ONE launch: for each entry
 * dw[param layout] += scale * ws[tap][co][ci], then ws is reset to zero.  `table` is a DEVICE
 * array; no two entries of one call may share a dw (the update is not atomic). */
typedef struct PgUnpackEntry {
  float *ws;
  float *dw;
  int Cin, Cout, Cin_p, Cout_p, taps, swap_io, flip, reserved;
  float scale, reserved2;
} PgUnpackEntry;
int pg_wgrad_unpack_multi(const PgUnpackEntry *table, int n, void *stream);

/* ---- PixelNorm + LeakyReLU derivatives: progan_modules.py:54-60,138 ------
 * y is the stored post-activation, r the stored per-pixel rsqrt.            */
/* stand-alone forward (PixelNorm.forward :58-60 + LeakyReLU :138) for layers wider than one
 * conv N tile (more than 256 output channels): y = lrelu(a * r), r = rsqrt(mean_c a^2 + 1e-8)
 * written to r[P] when use_pn (else r unused, y = lrelu(a)).  a and y may alias. */
int pg_pn_lrelu_fwd(const void *a, void *y, float *r, long long P, int C, float slope, int use_pn,
                    int dtype, void *stream);
/* pool_h/pool_w != 0: dy is the gradient of the 2x2-average-pooled activation
 * ([N,H/2,W/2,C]); the x1/4 expansion (avgpool backward) is fused.  colsum (nullable, fp32
 * [C], zero-initialised) += per-channel sum of da = bias gradient of the conv in front.
 * addend (nullable, [P,C]): a second gradient contribution to the same pre-activation (the
 * PixelNorm Hessian term of the GP sweep); da = backward(dy) + addend, colsum covers both. */
int pg_pn_lrelu_bwd(const void *dy, const void *y, const float *r, void *da,
                    long long P, int C, float slope, int use_pn, int pool_h, int pool_w,
                    float *colsum, const void *addend, int dtype, void *stream);
/* second order (WGAN-GP, train.py:146-151): given t = cotangent of da,
 * cot_dy = M Jpn t ;  cot_a = d/da <t, Jpn(a) M dy>                          */
int pg_pn_lrelu_bwd_bwd(const void *t, const void *dy, const void *y, const float *r,
                        void *cot_dy, void *cot_a, long long P, int C,
                        float slope, int use_pn, int pool_h, int pool_w, int dtype,
                        void *stream);
/* column sum over pixels (bias gradient): out[C] fp32 (zero-initialised) +=  */
int pg_colsum(const void *x, float *out, long long P, int C, int dtype, void *stream);

/* ---- 1x1 heads with 3 (or 1) image channels: from_rgb / to_rgb / linear ---
 * progan_modules.py:195-200,270-276,280.  img is NCHW fp32 [N,K,HW];
 * act is NHWC [N*HW, C]; logical weight w(c,k) = w[c*w_sc + k*w_sk].        */
int pg_pw_expand(const float *img, const float *w, const float *bias, void *act,
                 int N, long long HW, int K, int C, int w_sc, int w_sk,
                 float scale, int dtype, void *stream);
int pg_pw_reduce(const void *act, const float *w, const float *bias, float *img,
                 int N, long long HW, int K, int C, int w_sc, int w_sk,
                 float scale, int dtype, void *stream);
/* dw(c,k) += scale * sum_pix act[pix,c] img[k,pix]; dw zero-initialised.
 * dbias (may be NULL): dbias[c] += sum_pix act[pix,c] — the bias gradient of a
 * from_rgb layer, accumulated by the same pass over its output gradient      */
int pg_pw_wgrad(const void *act, const float *img, float *dw, float *dbias, int N,
                long long HW, int K, int C, int w_sc, int w_sk, float scale, int dtype,
                void *stream);
/* per-channel sum of an NCHW fp32 image tensor (to_rgb / linear bias grad)  */
int pg_img_chansum(const float *img, float *out, int N, long long HW, int K,
                   void *stream);

/* ---- resampling: F.interpolate bilinear x2 / x0.5, progan_modules.py:168,205,299 */
/* tensors are [N,H,W,C]; an NCHW fp32 image is passed as N*K planes with C=1 */
int pg_avgpool2(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream);
int pg_avgpool2_bwd(const void *dy, void *dx, int N, int H, int W, int C, int dtype, void *stream);
int pg_upsample2(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream);
int pg_upsample2_bwd(const void *dy, void *dx, int N, int H, int W, int C, int dtype, void *stream);

/* ---- fade-in blend: out = (1-alpha)*a + alpha*b (progan_modules.py:212,305)
 * alpha is read from DEVICE memory so a captured CUDA graph follows the
 * schedule.  pg_scale: out = (c0 + c1*alpha) * x.                            */
int pg_blend(const void *a, const void *b, void *out, long long n, const float *alpha_dev,
             int dtype, void *stream);
int pg_scale(const void *x, void *out, long long n, float c0, float c1,
             const float *alpha_dev, int dtype, void *stream);
int pg_tanh_fwd(const float *x, float *y, long long n, void *stream);
int pg_tanh_bwd(const float *dy, const float *y, float *dx, long long n, void *stream);

/* ---- minibatch-stddev: progan_modules.py:289-293 ---------------------------
 * x:[N,F] (F = 16*C features of the 4x4 map, NHWC so feature f = pos*C + c),
 * out:[N,16,Cp] = x with channel C set to mean_f sigma_f and channels
 * C+1..Cp-1 zero.  stats: fp32 workspace of 4*F + 64 floats written by fwd
 * (mu, sigma per feature) and reused by bwd / bwd_bwd.                        */
int pg_mbstd_fwd(const void *x, void *out, float *stats, int N, int C, int Cp, int dtype,
                 void *stream);
int pg_mbstd_bwd(const void *dout, const void *x, const float *stats, void *dx, int N, int C,
                 int Cp, int dtype, void *stream);
int pg_mbstd_bwd_bwd(const void *t, const void *dout, const void *x, float *stats,
                     void *cot_dout, void *cot_x, int N, int C, int Cp, int dtype,
                     void *stream);

/* ---- losses: train.py:126-139 (critic, with the 0.001 drift term) / :162-167 (generator) ----
 * d: critic outputs, the n_real real samples first, then n_fake fake ones (fp32).
 *   n_real > 0:  L = -(mean d_r - drift*mean d_r^2) + mean d_f;  metric[0] += mean d_r - drift*mean d_r^2 - mean d_f
 *   n_real = 0:  L = -mean d;                                     metric[0] += L
 * seed[n] = dL/dd[n] (what the reference passes implicitly through .backward(one/mone)). */
int pg_wgan_loss(const float *d, float *seed, float *metric, int n_real, int n_fake, float drift,
                 void *stream);

/* ---- WGAN-GP pieces: train.py:142-150 ---------------------------------------*/
/* x_hat = eps[n]*real + (1-eps[n])*fake  (all fp32 [N,D])                    */
int pg_interp_xhat(const float *real, const float *fake, const float *eps, float *out,
                   int N, long long D, void *stream);
/* norms[n] = ||g[n,:]||_2 ; gp = lambda/N * sum (norms-1)^2                   */
int pg_gp_fwd(const float *g, float *norms, float *gp, int N, long long D, float lambda,
              void *stream);
/* v[n,:] = upstream * 2*lambda/N * (norm-1)/norm * g[n,:]                      */
int pg_gp_bwd(const float *g, const float *norms, const float *upstream, float *v, int N,
              long long D, float lambda, void *stream);

/* ---- optimiser: Adam(beta1,beta2) over a flat bucket ----
 * train.py:256-257.  m may be NULL when beta1 == 0.  step_dev holds the
 * 1-based step count as float (bias correction computed on device).          */
int pg_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr,
                 float beta1, float beta2, float eps, const float *step_dev,
                 float grad_scale, void *stream);
/* multi-tensor form: chunks is an int4 table {start, length, segment, 0} over the flat
 * bucket, steps_dev[segment] the per-parameter-group step count (torch.optim.Adam keeps
 * one step per parameter and skips parameters without a gradient).                  */
int pg_adam_multi(float *p, const float *g, float *m, float *v, const void *chunks,
                  int nchunks, const float *steps_dev, float lr, float beta1, float beta2,
                  float eps, float grad_scale, void *stream);
/* ema = decay*ema + (1-decay)*p   (accumulate(), train.py:17-22)              */
int pg_ema(float *ema, const float *p, long long n, float decay, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PROGAN_B200_H */
