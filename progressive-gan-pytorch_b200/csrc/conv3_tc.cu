// 3x3 pad-1 implicit-GEMM convolution, second generation ("column-halo") kernel.
//
// Measured on B200 (profiles/README.md): the first-generation kernel (conv_tc.cu) is bound
// by L2->SM request throughput (~4.3 cycles per 128-byte line per SM, independent of the
// pipeline depth), because every K block re-reads a shifted copy of the activation tile
// (9x) and a weight tile.  This kernel cuts the lines fetched per MAC:
//
//  * pixel tile = 8 (w) x 16 (h); for every 64-channel block only THREE activation boxes are
//    fetched, one per horizontal tap offset dw, each 8 x 18 pixels (one halo row above and
//    below).  Because a tile row is exactly 8 pixels x 128 B = 1024 B, the three vertical
//    taps dh are plain 1024-byte-aligned start-address offsets of the same smem tile, so the
//    UMMA descriptors stay canonical (SWIZZLE_128B, SBO = 1024).  A traffic: 3*18 vs 9*16 rows.
//  * weights: if the whole packed matrix [Cout][9][Cin] fits next to the pipeline it is loaded
//    into smem ONCE per CTA ("resident"); otherwise the three taps of the current dw are
//    streamed with the activation unit and TWO pixel tiles share them (M = 2 x 128 per weight
//    tile, two accumulators in TMEM).
//
// Everything else follows conv_tc.cu: tcgen05.mma cta_group::1 kind::f16 (M128 x N=Cout x K16),
// double-buffered fp32 accumulators in TMEM, TMA producer warp / single-thread MMA issuer /
// 4 epilogue warps with the fused scale+bias+PixelNorm+LeakyReLU epilogue and TMA stores.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pg {
namespace tc {

struct Conv3Params {
  int N, H, W, Cin, Cout;
  int tiles_w, tiles_h, num_tiles;     // 8x16 pixel tiles
  int MT;                              // pixel tiles sharing one weight unit (1 or 2)
  int num_super;                       // num_tiles / MT
  int BK, ncb;                         // channel block (64/32) and count
  int resident;                        // weights resident in smem
  int units;                           // pipeline depth (units)
  int a_tile_bytes;                    // one 8x18 activation box
  int b_tap_bytes;                     // one weight box: Cout x BK
  int unit_bytes;
  int out_chunk, tmem_cols;
  int epi;
  float scale, slope;
  const float *bias;
  float *r_out;
  int dbg;                             // experiment knobs (PG_DBG): 1 no store, 2 no pass-2, 4 no epilogue
};

constexpr int kC3Threads = 384;   // 4 control warps + 8 epilogue warps

// 64-bit smem descriptor from a precomputed low word (start>>4 | LBO) and a constant high word
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}

template <int BK, int MT, bool RES, int COUT>
__global__ void __launch_bounds__(kC3Threads, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_y, const Conv3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wres_bytes = p.resident ? (uint32_t)(9 * p.ncb * p.b_tap_bytes) : 0u;
  const uint32_t smem_w = base;                                  // resident weights
  const uint32_t smem_u0 = base + wres_bytes;                    // unit ring
  const uint32_t smem_out = smem_u0 + (uint32_t)(p.units * p.unit_bytes);
  const uint32_t out_bytes = 128u * (uint32_t)p.Cout * 2u;
  const uint32_t bar_base = smem_out + out_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.units + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.units + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.units + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (uint32_t)(2 * p.units + 4);
  const uint32_t tmem_slot = wfull_bar + 8u;
  const uint32_t bias_s = (tmem_slot + 16u + 15u) & ~15u;      // float4-aligned
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gbase + (tmem_slot - base));
  float *bias_ptr = reinterpret_cast<float *>(gbase + (bias_s - base));
  float *ss_buf = bias_ptr + 128;                    // [2][128] partial sums of squares
  uint8_t *out_ptr = gbase + (smem_out - base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.units; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int c = threadIdx.x; c < p.Cout; c += kC3Threads) bias_ptr[c] = p.bias ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr int acc_stride = MT * COUT;       // TMEM columns per accumulator stage

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (RES) {
        mbar_expect_tx(wfull_bar, wres_bytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int cb = 0; cb < p.ncb; ++cb)
            tma_load_2d(smem_w + (uint32_t)((tap * p.ncb + cb) * p.b_tap_bytes), &tmap_w, wfull_bar,
                        tap * p.Cin + cb * p.BK, 0);
      }
      int u = 0;
      uint32_t phase = 0;
      for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
        for (int cb = 0; cb < p.ncb; ++cb) {
          for (int dwi = 0; dwi < 3; ++dwi) {
            mbar_wait(empty_bar(u), phase ^ 1u);
            if (p.dbg & 8) {            // experiment: no loads at all (timing of MMA+epilogue only)
              mbar_arrive(full_bar(u));
              if (++u == p.units) { u = 0; phase ^= 1u; }
              continue;
            }
            mbar_expect_tx(full_bar(u), (uint32_t)p.unit_bytes);
            const uint32_t su = smem_u0 + (uint32_t)(u * p.unit_bytes);
            for (int mt = 0; mt < MT; ++mt) {
              const int tile = st * MT + mt;
              const int tw = tile % p.tiles_w;
              const int th = (tile / p.tiles_w) % p.tiles_h;
              const int n = tile / (p.tiles_w * p.tiles_h);
              tma_load_4d(su + (uint32_t)(mt * p.a_tile_bytes), &tmap_x, full_bar(u), cb * p.BK,
                          tw * 8 + dwi - 1, th * 16 - 1, n);
            }
            if (!RES) {
              const uint32_t sb = su + (uint32_t)(MT * p.a_tile_bytes);
              for (int dh = 0; dh < 3; ++dh)
                tma_load_2d(sb + (uint32_t)(dh * p.b_tap_bytes), &tmap_w, full_bar(u),
                            (dh * 3 + dwi) * p.Cin + cb * p.BK, 0);
            }
            if (++u == p.units) {
              u = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the (warp-uniform) loop so that addresses and descriptors live in
    // uniform registers; only the tcgen05 instructions are predicated to lane 0.  Descriptors
    // are a constant high word + (start>>4) low word advanced by compile-time offsets: a few
    // instructions per MMA instead of a dependent 64-bit build (measured: the issue loop,
    // not the memory system, bounded the first version at ~180 cycles per MMA).
    constexpr uint32_t row_bytes = BK * 2u;                          // 128 or 64
    constexpr uint32_t layout = row_bytes == 128 ? 2u : 4u;
    constexpr uint32_t sbo = 8u * row_bytes;                          // one 8-pixel tile row
    constexpr int nk = BK / 16;
    constexpr uint32_t a_tile = 18u * 8u * row_bytes;
    constexpr uint32_t desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    constexpr uint32_t lbo_lo = 1u << 16;                             // LBO = 16 B (unused, canonical)
    const uint32_t idesc = make_idesc_bf16(128, COUT, 0, 0);
    const uint32_t b_tap16 = (uint32_t)p.b_tap_bytes >> 4;
    const uint32_t b_dh16 = RES ? 3u * (uint32_t)p.ncb * b_tap16 : b_tap16;   // next dh's weight box
    if (RES) {
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
    }
    int u = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(acc * acc_stride);
      for (int cb = 0; cb < p.ncb; ++cb) {
#pragma unroll
        for (int dwi = 0; dwi < 3; ++dwi) {
          mbar_wait(full_bar(u), phase);
          tc_fence_after();
          const uint32_t su = smem_u0 + (uint32_t)(u * p.unit_bytes);
          const uint32_t a_lo = (su >> 4) | lbo_lo;
          const uint32_t b_addr = RES ? smem_w + (uint32_t)((dwi * p.ncb + cb) * p.b_tap_bytes)
                                      : su + MT * a_tile;
          const uint32_t b_lo = (b_addr >> 4) | lbo_lo;
          const uint32_t first = (uint32_t)((cb | dwi) != 0);
          if (elect_one_sync()) {
            if (!(p.dbg & 16))          // experiment: no MMAs (timing of loads + epilogue only)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
                for (int k = 0; k < nk; ++k) {
                  const uint32_t al = a_lo + ((mt * a_tile + dh * sbo + k * 32u) >> 4);
                  const uint32_t bl = b_lo + dh * b_dh16 + ((k * 32u) >> 4);
                  umma_bf16(d_base + (uint32_t)(mt * COUT), desc64(al, desc_hi), desc64(bl, desc_hi),
                            idesc, (dh | k) ? 1u : first);
                }
              }
            }
            umma_commit(empty_bar(u));
            if (cb == p.ncb - 1 && dwi == 2) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++u == p.units) {
            u = 0;
            phase ^= 1u;
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // Two warps per TMEM lane quadrant, each owning half of the channel columns of its 32
    // pixels: one tcgen05.ld pass, values stay in registers; the PixelNorm sum of squares is
    // completed through a 1 KB smem exchange between the two halves.
    constexpr int CPT = COUT / 2;                 // columns per thread: 16 / 32 / 64
    constexpr int out_chunk = (COUT % 64 == 0) ? 64 : 32;
    constexpr int chunk_rows_bytes = out_chunk * 2;
    constexpr int swz_bits = chunk_rows_bytes == 128 ? 3 : 2;
    constexpr int n_chunks = COUT / out_chunk;
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;                // tile row: pixel (hl = row/8, wl = row%8)
    const int et = threadIdx.x - 128;             // 0..255
    const int col0 = half * CPT;
    const float invC = 1.f / (float)COUT;
    const float scale = p.scale, slope = p.slope;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int tile = st * MT + mt;
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                (uint32_t)(acc * acc_stride + mt * COUT + col0);
        uint32_t vr[CPT];
        tmem_ld<CPT>(t_addr, vr);
        tmem_ld_wait();
        if (mt == MT - 1) {            // accumulator stage fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
        float v[CPT];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4 *>(bias_ptr + col0 + j);
          v[j] = fmaf(__uint_as_float(vr[j]), scale, b4.x);
          v[j + 1] = fmaf(__uint_as_float(vr[j + 1]), scale, b4.y);
          v[j + 2] = fmaf(__uint_as_float(vr[j + 2]), scale, b4.z);
          v[j + 3] = fmaf(__uint_as_float(vr[j + 3]), scale, b4.w);
          ss = fmaf(v[j], v[j], ss);
          ss = fmaf(v[j + 1], v[j + 1], ss);
          ss = fmaf(v[j + 2], v[j + 2], ss);
          ss = fmaf(v[j + 3], v[j + 3], ss);
        }
        float r = 1.f;
        if (p.epi == PG_EPI_PN_LRELU) {
          ss_buf[half * 128 + row] = ss;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          r = rsqrtf((ss_buf[row] + ss_buf[128 + row]) * invC + 1e-8f);
        }
        if (et == 0) tma_store_wait_read0();       // staging buffer free again?
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          constexpr int chunk = 0;  (void)chunk;
          const int c_abs = col0;                                  // first column of this thread
          const int chunk_i = c_abs / out_chunk;
          const int cin_chunk = c_abs - chunk_i * out_chunk;
          uint8_t *tile_base = out_ptr + (size_t)chunk_i * 128 * chunk_rows_bytes;
#pragma unroll
          for (int i = 0; i < CPT / 8; ++i) {                      // 8 channels = 16 bytes per store
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a0 = v[i * 8 + 2 * e] * r, a1 = v[i * 8 + 2 * e + 1] * r;
              if (p.epi != PG_EPI_LINEAR) {
                a0 = a0 > 0.f ? a0 : a0 * slope;
                a1 = a1 > 0.f ? a1 : a1 * slope;
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              pk[e] = *reinterpret_cast<uint32_t *>(&h);
            }
            // CPT <= out_chunk except COUT=128 (CPT=64=out_chunk): a thread never straddles chunks
            const uint32_t off = (uint32_t)row * (uint32_t)chunk_rows_bytes +
                                 (uint32_t)cin_chunk * 2u + (uint32_t)i * 16u;
            *reinterpret_cast<uint4 *>(tile_base + swz(off, swz_bits)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        if (p.epi == PG_EPI_PN_LRELU && half == 0)
          p.r_out[((long long)n * p.H + (h0 + (row >> 3))) * p.W + w0 + (row & 7)] = r;
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && !(p.dbg & 1)) {
#pragma unroll
          for (int ch = 0; ch < n_chunks; ++ch)
            tma_store_4d(&tmap_y, smem_out + (uint32_t)ch * 128u * (uint32_t)chunk_rows_bytes,
                         ch * out_chunk, w0, h0, n);
          tma_store_commit();
        }
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

// Returns PG_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible,
// so pg_conv_tc can fall through to the first-generation kernel.
int conv3_tc_launch(const void *x, const void *wp, const float *bias, void *y, float *r_out, int N,
                    int H, int W, int Cin, int Cout, float scale, int epi, float slope,
                    cudaStream_t stream) {
  if (const char *e = getenv("PG_CONV_V2"))
    if (atoi(e) == 0) return PG_ERR_UNSUPPORTED;
  if (H % 16 || H < 32 || W % 8 || !(Cout == 32 || Cout == 64 || Cout == 128) || Cin % 32)
    return PG_ERR_UNSUPPORTED;
  tc::Conv3Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * N;
  p.BK = (Cin % 64 == 0) ? 64 : 32;
  p.ncb = Cin / p.BK;
  p.a_tile_bytes = 18 * 8 * p.BK * 2;
  p.b_tap_bytes = Cout * p.BK * 2;
  p.out_chunk = (Cout % 64 == 0) ? 64 : 32;
  p.epi = epi; p.scale = scale; p.slope = slope; p.bias = bias; p.r_out = r_out;
  p.dbg = 0;
  if (const char *e = getenv("PG_DBG")) p.dbg = atoi(e);
  const int out_bytes = 128 * Cout * 2;
  const int misc = 1024 + 8 * (2 * 12 + 5) + 16 + 128 * 4 + 2 * 128 * 4 + 64;
  const int budget = 227 * 1024 - out_bytes - misc;
  const int wres = 9 * p.ncb * p.b_tap_bytes;
  // choose (resident?, MT): fewest L2 lines per pixel tile subject to smem / TMEM limits.
  // lines/tile = activation rows (3 boxes x 144 rows per channel block) + streamed weight rows
  // (9 x Cout per channel block, shared by MT tiles).  Ties go to the larger MT (per-tile
  // barrier/epilogue overheads amortise).
  int force_res = -1, force_mt = -1;
  if (const char *e = getenv("PG_CONV_V2_MODE")) force_res = atoi(e);   // 0 stream, 1 resident
  if (const char *e = getenv("PG_CONV_V2_MT")) force_mt = atoi(e);
  double best = 1e30;
  p.units = 0;
  for (int res = 1; res >= 0; --res) {
    if (force_res >= 0 && res != force_res) continue;
    for (int mt = 4; mt >= 1; mt >>= 1) {
      if (force_mt > 0 && mt != force_mt) continue;
      if (2 * mt * Cout > 512 || p.num_tiles % mt) continue;
      const int unit = mt * p.a_tile_bytes + (res ? 0 : 3 * p.b_tap_bytes);
      const int avail = budget - (res ? wres : 0);
      if (avail < 2 * unit) continue;
      int units = avail / unit;
      if (units > 12) units = 12;
      const double lines = 3.0 * p.ncb * 144 + (res ? 0.0 : 9.0 * p.ncb * Cout / mt);
      const double score = lines - 1e-3 * mt;
      if (score < best) {
        best = score;
        p.resident = res; p.MT = mt; p.unit_bytes = unit; p.units = units;
      }
    }
  }
  if (p.units < 2) return PG_ERR_UNSUPPORTED;
  p.num_super = p.num_tiles / p.MT;
  int cols = 2 * p.MT * Cout;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const size_t smem = (size_t)(p.resident ? wres : 0) + (size_t)p.units * p.unit_bytes + out_bytes + misc;

  CUtensorMap tx, tw_, ty;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.BK, 8u, 18u, 1u};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, p.BK * 2, "pg_conv_tc/v2(x)")) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {(uint32_t)p.BK, (uint32_t)Cout};
    if (int rc = make_tmap_bf16(&tw_, wp, 2, dims, str, box, p.BK * 2, "pg_conv_tc/v2(w)")) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {(uint32_t)p.out_chunk, 8u, 16u, 1u};
    if (int rc = make_tmap_bf16(&ty, y, 4, dims, str, box, p.out_chunk * 2, "pg_conv_tc/v2(y)")) return rc;
  }
  int grid = p.num_super < sm_count() ? p.num_super : sm_count();
  cudaError_t e = cudaSuccess;
#define PG_C3_LAUNCH(BK_, MT_, RES_, CO_)                                                        \
  {                                                                                              \
    static bool attr_set = false;                                                                \
    if (!attr_set) {                                                                             \
      e = cudaFuncSetAttribute(tc::conv3_tc_kernel<BK_, MT_, RES_, CO_>,                         \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);         \
      attr_set = (e == cudaSuccess);                                                             \
    }                                                                                            \
    if (e == cudaSuccess)                                                                        \
      tc::conv3_tc_kernel<BK_, MT_, RES_, CO_><<<grid, tc::kC3Threads, smem, stream>>>(tx, tw_, ty, p); \
  }
#define PG_C3_RES(BK_, MT_, CO_)                                                                 \
  {                                                                                              \
    if (p.resident) PG_C3_LAUNCH(BK_, MT_, true, CO_)                                            \
    else PG_C3_LAUNCH(BK_, MT_, false, CO_)                                                      \
  }
#define PG_C3_MODE(BK_, CO_)                                                                     \
  {                                                                                              \
    if (p.MT == 1) PG_C3_RES(BK_, 1, CO_)                                                        \
    else if (p.MT == 2) PG_C3_RES(BK_, 2, CO_)                                                   \
    else if (CO_ <= 64) PG_C3_RES(BK_, (CO_ <= 64 ? 4 : 2), CO_)                                 \
  }
  if (p.BK == 64) {
    if (Cout == 128) PG_C3_MODE(64, 128)
    else if (Cout == 64) PG_C3_MODE(64, 64)
    else PG_C3_MODE(64, 32)
  } else {
    if (Cout == 128) PG_C3_MODE(32, 128)
    else if (Cout == 64) PG_C3_MODE(32, 64)
    else PG_C3_MODE(32, 32)
  }
#undef PG_C3_RES
#undef PG_C3_MODE
#undef PG_C3_LAUNCH
  if (e != cudaSuccess) {
    set_error("pg_conv_tc/v2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pg_conv_tc/v2: CUDA launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace pg
