// 3x3 pad-1 (and 1x1) implicit-GEMM convolution on the 5th-gen tensor cores.
//
// Replaces aten::convolution (cuDNN fprop / dgrad) for the EqualConv2d 3x3 layers of
// ConvBlock (reference progan_modules.py:63-73,120-148) and their data-gradients (the
// data-gradient of a 3x3 pad-1 conv is the same conv with flipped/transposed weights, so
// one kernel serves forward, dgrad and the GP tangent conv).  The equalized-LR scale
// (:22-27), bias, PixelNorm (:54-60) and LeakyReLU(0.2) are fused into the epilogue.
//
// GEMM view: M = N*H*W output pixels (tile = 128 pixels = a bw x bh x bn spatial box),
//            N = Cout (whole channel vector in one tile -> PixelNorm is thread-local),
//            K = taps*Cin, walked as (tap, 64- or 32-channel block).
// A tiles : TMA tiled-mode 4-D boxes over the NHWC activation, shifted by the tap offset;
//           out-of-bounds coordinates are zero-filled by the TMA unit == conv padding.
// B tiles : TMA 2-D boxes over the K-major packed weight matrix [Cout][taps*Cin].
// MMA     : tcgen05.mma.cta_group::1.kind::f16, M=128, N=Cout, K=16, fp32 accumulators in
//           TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Roles   : warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//           warps 4-7 = epilogue (tcgen05.ld -> scale/bias/PN/LReLU -> bf16 -> swizzled
//           smem -> TMA store).  Persistent: grid = min(#tiles, #SMs).
#include "tc_common.cuh"
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <mutex>
#include <stdlib.h>

namespace pg {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

// Descriptor cache ("caches (TMA descriptors keyed by ptr/shape) are per-process,
// mutex-guarded"): a tensor map is a pure function of (base, dims, strides, box, swizzle), and
// the caching allocator hands the same buffers back every iteration, so the eager path encodes
// each descriptor once instead of three or four driver calls per launch.
namespace {
struct TmapKey {
  const void *base;
  int rank, swizzle;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  bool operator==(const TmapKey &o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey &k) const {
    const uint64_t *w = reinterpret_cast<const uint64_t *>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
}  // namespace

int make_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims,
                   const uint64_t *strides_bytes, const uint32_t *box, int swizzle_bytes,
                   const char *what) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.swizzle = swizzle_bytes;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return PG_OK;
    }
  }
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)", what);
    return PG_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, (void *)base, gdim,
                   gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
    return PG_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();      // bound the table (shape sweeps)
    g_tmap_cache.emplace(key, *out);
  }
  return PG_OK;
}

namespace tc {

struct ConvTcParams {
  int N, H, W, Cin, Cout, taps;   // Cout = channels of ONE N tile (<= 256)
  int n_tiles;                    // N tiles (Cout_total / Cout); GEMM mode for wide outputs
  int bw, bh, bn;            // spatial box of one 128-pixel tile
  int tiles_w, tiles_h, tiles_n, num_tiles;
  int BK, kb_per_tap, num_kb;  // K block (channels) and counts
  int stages;
  int a_bytes, b_bytes;      // per-stage operand tile sizes
  int out_chunk;             // channels per output TMA box (<=64)
  int tmem_cols;             // allocated TMEM columns (2 accumulator stages)
  int epi;
  float scale, slope;
  const float *bias;
  int bias_mod;              // bias[c % bias_mod]
  int bias_per_tile;         // N tile nt uses bias[(nt * Cout + c) % bias_mod] (bias_mod > Cout)
  float *r_out;
};

constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_y, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (base is 1024-aligned by the launch: dynamic smem starts aligned; enforce anyway)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)(p.a_bytes + p.b_bytes);
  const uint32_t smem_a0 = base;
  const uint32_t smem_out = base + (uint32_t)p.stages * stage_bytes;
  const uint32_t out_bytes = 128u * (uint32_t)p.Cout * 2u;
  const uint32_t bar_base = smem_out + out_bytes;           // 8-byte aligned (multiple of 1024)
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * p.stages + 4);
  const uint32_t bias_s = tmem_slot + 16u;
  // generic pointers for plain smem accesses
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(gbase + (tmem_slot - base));
  float *bias_ptr = reinterpret_cast<float *>(gbase + (bias_s - base));
  uint8_t *out_ptr = gbase + (smem_out - base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // everything above touches only this CTA's shared memory / TMEM: it overlaps the tail of the
  // previous kernel; from here on the predecessors' results are read
  pg::grid_dep_sync();
  for (int c = threadIdx.x; c < p.Cout; c += kThreads)
    bias_ptr[c] = p.bias ? p.bias[c % p.bias_mod] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;          // N tile fastest: neighbours share the A tile in L2
        const int pt = tile / p.n_tiles;
        const int tw = pt % p.tiles_w;
        const int th = (pt / p.tiles_w) % p.tiles_h;
        const int tn = pt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          int dh = 0, dw = 0;
          if (p.taps == 9) {
            dh = tap / 3 - 1;
            dw = tap % 3 - 1;
          }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), stage_bytes);
          const uint32_t sa = smem_a0 + (uint32_t)stage * stage_bytes;
          tma_load_4d(sa, &tmap_x, full_bar(stage), cb * p.BK, w0 + dw, h0 + dh, n0);
          tma_load_2d(sa + (uint32_t)p.a_bytes, &tmap_w, full_bar(stage),
                      tap * p.Cin + cb * p.BK, nt * p.Cout);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop; tcgen05 ops issued by the elected lane; precomputed descriptors
    const uint32_t idesc = make_idesc_bf16(128, p.Cout, 0, 0);
    const uint32_t row_bytes = (uint32_t)p.BK * 2u;            // 128 or 64
    const uint32_t layout = row_bytes == 128 ? 2u : 4u;         // SWIZZLE_128B / SWIZZLE_64B
    const uint32_t desc_hi = (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    const int nk = p.BK / 16;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * (p.tmem_cols / 2));
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_a0 + (uint32_t)stage * stage_bytes;
        const uint32_t a_lo = (sa >> 4) | (1u << 16);
        const uint32_t b_lo = ((sa + (uint32_t)p.a_bytes) >> 4) | (1u << 16);
        if (elect_one_sync()) {
          if (nk == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k),
                        ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k), idesc,
                        (uint32_t)((kb | k) != 0));
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k),
                        ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k), idesc,
                        (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));               // frees the smem slot when MMAs retire
          if (kb == p.num_kb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                     // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;              // tile row == TMEM lane
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..127
    const int chunk_rows_bytes = p.out_chunk * 2;  // 128 or 64
    const int swz_bits = chunk_rows_bytes == 128 ? 3 : 2;
    const int n_chunks = p.Cout / p.out_chunk;
    const float invC = 1.f / (float)p.Cout;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int pt = tile / p.n_tiles;
      const int tw = pt % p.tiles_w;
      const int th = (pt / p.tiles_w) % p.tiles_h;
      const int tn = pt / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
      if (p.bias_per_tile) {      // wide outputs: every N tile has its own slice of the bias
        asm volatile("bar.sync 1, 128;" ::: "memory");     // previous tile's reads are done
        for (int c = et; c < p.Cout; c += 128) bias_ptr[c] = p.bias[(nt * p.Cout + c) % p.bias_mod];
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr =
          tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * (p.tmem_cols / 2));
      float r = 1.f;
      if (p.epi == PG_EPI_PN_LRELU) {
        float ss = 0.f;
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = __uint_as_float(v[j]) * p.scale + bias_ptr[c0 + j];
            ss = fmaf(a, a, ss);
          }
        }
        r = rsqrtf(ss * invC + 1e-8f);
      }
      // the previous tile's TMA store must have finished reading the staging buffer
      if (et == 0) tma_store_wait_read0();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a0 = (__uint_as_float(v[j]) * p.scale + bias_ptr[c0 + j]) * r;
          float a1 = (__uint_as_float(v[j + 1]) * p.scale + bias_ptr[c0 + j + 1]) * r;
          if (p.epi != PG_EPI_LINEAR) {
            a0 = a0 > 0.f ? a0 : a0 * p.slope;
            a1 = a1 > 0.f ? a1 : a1 * p.slope;
          }
          __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
          packed[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
        }
        // 32 channels = 64 bytes = four 16-byte chunks of this row, swizzled like the TMA box
        const int chunk = c0 / p.out_chunk;
        const int cin_chunk = c0 - chunk * p.out_chunk;  // channel offset inside the chunk
        uint8_t *tile_base = out_ptr + (size_t)chunk * 128 * chunk_rows_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)row * (uint32_t)chunk_rows_bytes +
                               (uint32_t)cin_chunk * 2u + (uint32_t)i * 16u;
          *reinterpret_cast<uint4 *>(tile_base + swz(off, swz_bits)) =
              make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
      }
      // accumulator stage is drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (p.epi == PG_EPI_PN_LRELU) {
        const int wl = row % p.bw, hl = (row / p.bw) % p.bh, nl = row / (p.bw * p.bh);
        const int w = w0 + wl, h = h0 + hl, n = n0 + nl;
        if (w < p.W && h < p.H && n < p.N)
          p.r_out[(((long long)n * p.H + h) * p.W + w) * p.n_tiles + nt] = r;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        for (int ch = 0; ch < n_chunks; ++ch)
          tma_store_4d(&tmap_y, smem_out + (uint32_t)ch * 128u * (uint32_t)chunk_rows_bytes,
                       nt * p.Cout + ch * p.out_chunk, w0, h0, n0);
        tma_store_commit();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace tc
}  // namespace pg

using namespace pg;

static int pow2_ge(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

extern "C" int pg_conv_tc(const void *x, const void *wp, const float *bias, void *y, float *r_out,
                          int N, int H, int W, int Cin, int Cout_total, int Cout_tile, int taps,
                          int bias_mod, float scale, int epi, float slope, void *y_pool,
                          void *stream) {
  PG_CHECK_ARG(x && wp && y, "pg_conv_tc: null pointer");
  PG_CHECK_ARG(taps == 9 || taps == 1, "pg_conv_tc: taps must be 9 (3x3 pad 1) or 1");
  PG_CHECK_ARG(N > 0 && H > 0 && W > 0, "pg_conv_tc: bad dims");
  PG_CHECK_ARG(Cin % 32 == 0 && Cin > 0, "pg_conv_tc: Cin %% 32 != 0 (Cin=%d)", Cin);
  const int Cout = Cout_tile;
  PG_CHECK_ARG(Cout % 32 == 0 && Cout >= 32 && Cout <= 256,
               "pg_conv_tc: N tile must be a multiple of 32 in [32,256] (%d)", Cout);
  PG_CHECK_ARG(Cout_total % Cout == 0, "pg_conv_tc: Cout_total %d not a multiple of the tile %d",
               Cout_total, Cout);
  const int n_tiles = Cout_total / Cout;
  const bool bias_per_tile = bias && n_tiles > 1 && bias_mod > Cout && bias_mod % Cout == 0;
  PG_CHECK_ARG(!bias || bias_per_tile ||
                   (bias_mod > 0 && (n_tiles == 1 ? bias_mod == Cout : Cout % bias_mod == 0)),
               "pg_conv_tc: bias_mod %d incompatible with the N tiling", bias_mod);
  PG_CHECK_ARG(epi != PG_EPI_PN_LRELU || r_out, "pg_conv_tc: PN epilogue needs r_out");
  PG_CHECK_ARG(epi != PG_EPI_PN_LRELU || n_tiles == 1 || (taps == 1 && !bias_per_tile),
               "pg_conv_tc: PixelNorm over more channels than one N tile (<= 256): use "
               "PG_EPI_LINEAR + pg_pn_lrelu_fwd (Cout_total %d, tile %d)", Cout_total, Cout);
  PG_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "pg_conv_tc: pointers must be 16-byte aligned");
  if (taps == 9 && n_tiles == 1) {   // the single-halo-box kernel where the shape allows
    const int rc4 = conv4_tc_launch(x, wp, bias, y, r_out, N, H, W, Cin, Cout, scale, epi, slope,
                                    (cudaStream_t)stream, nullptr, nullptr, nullptr, 0, y_pool);
    if (rc4 != PG_ERR_UNSUPPORTED) return rc4;
  }
  PG_CHECK_ARG(!y_pool, "pg_conv_tc: the fused 2x2 pool needs a 3x3 conv with H %% 16 == 0, W %% 8 == 0, "
                        "Cout in {32,64,128} (H=%d W=%d Cout=%d taps=%d)", H, W, Cout, taps);
  tc::ConvTcParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps; p.n_tiles = n_tiles;
  p.bw = W < 16 ? W : 16;
  {
    int rem = 128 / p.bw;
    p.bh = H < rem ? H : rem;
    p.bn = 128 / (p.bw * p.bh);
  }
  PG_CHECK_ARG(p.bw * p.bh * p.bn == 128 && p.bn <= 256,
               "pg_conv_tc: cannot tile %dx%d into 128-pixel boxes", H, W);
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * n_tiles;
  p.BK = (Cin % 64 == 0) ? 64 : 32;
  p.kb_per_tap = Cin / p.BK;
  p.num_kb = taps * p.kb_per_tap;
  p.a_bytes = 128 * p.BK * 2;
  p.b_bytes = Cout * p.BK * 2;
  p.out_chunk = (Cout % 64 == 0) ? 64 : 32;
  p.tmem_cols = pow2_ge(2 * Cout);
  p.epi = epi; p.scale = scale; p.slope = slope; p.bias = bias; p.bias_mod = bias_mod > 0 ? bias_mod : 1;
  p.r_out = r_out;
  p.bias_per_tile = bias_per_tile ? 1 : 0;
  const int out_bytes = 128 * Cout * 2;
  const int misc = 1024 /*align slack*/ + 8 * (2 * 8 + 4) + 16 + Cout * 4 + 64;
  const int budget = 227 * 1024;
  int stages = (budget - out_bytes - misc) / (p.a_bytes + p.b_bytes);
  if (stages > 8) stages = 8;
  PG_CHECK_ARG(stages >= 2, "pg_conv_tc: not enough shared memory for the pipeline");
  p.stages = stages;
  const size_t smem = (size_t)stages * (p.a_bytes + p.b_bytes) + out_bytes + misc;

  CUtensorMap tx, tw_, ty;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, p.BK * 2, "pg_conv_tc(x)")) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)taps * Cin, (uint64_t)Cout_total};
    uint64_t str[1] = {(uint64_t)taps * Cin * 2};
    uint32_t box[2] = {(uint32_t)p.BK, (uint32_t)Cout};
    if (int rc = make_tmap_bf16(&tw_, wp, 2, dims, str, box, p.BK * 2, "pg_conv_tc(w)")) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout_total, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout_total * 2, (uint64_t)W * Cout_total * 2,
                       (uint64_t)H * W * Cout_total * 2};
    uint32_t box[4] = {(uint32_t)p.out_chunk, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (int rc = make_tmap_bf16(&ty, y, 4, dims, str, box, p.out_chunk * 2, "pg_conv_tc(y)"))
      return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::conv_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("pg_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  pg::launcher(tc::conv_tc_kernel, grid, tc::kThreads, smem, (cudaStream_t)stream, p.num_tiles)(tx, tw_, ty, p);
  PG_CHECK_LAUNCH("pg_conv_tc");
}

// data-gradient conv with the activation backward of the PREVIOUS layer fused into the epilogue
extern "C" int pg_conv_tc_actbwd(const void *x, const void *wp, void *da, int N, int H, int W,
                                 int Cin, int Cout, float scale, const void *y_prev,
                                 const float *r_prev, float slope, int use_pn, float *colsum,
                                 void *stream) {
  PG_CHECK_ARG(x && wp && da && y_prev, "pg_conv_tc_actbwd: null pointer");
  PG_CHECK_ARG(!use_pn || r_prev, "pg_conv_tc_actbwd: PixelNorm form needs r_prev");
  PG_CHECK_ARG(slope > 0.f, "pg_conv_tc_actbwd: slope must be > 0");
  PG_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)da & 15) == 0 &&
                   ((uintptr_t)y_prev & 15) == 0,
               "pg_conv_tc_actbwd: pointers must be 16-byte aligned");
  const int rc = conv4_tc_launch(x, wp, nullptr, da, nullptr, N, H, W, Cin, Cout, scale,
                                 PG_EPI_LINEAR, slope, (cudaStream_t)stream, y_prev, r_prev, colsum,
                                 use_pn);
  if (rc == PG_ERR_UNSUPPORTED)
    set_error("pg_conv_tc_actbwd: shape N=%d H=%d W=%d Cin=%d Cout=%d is not served by the fused "
              "kernel (use pg_conv_tc + pg_pn_lrelu_bwd)", N, H, W, Cin, Cout);
  return rc;
}
