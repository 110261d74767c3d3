// 3x3 pad-1 implicit-GEMM convolution, third generation ("single halo box") kernel.
//
// conv3_tc.cu is bound by L2->SM bandwidth (~30-40 B/cycle/SM measured): per 128-pixel tile
// and 64-channel block it fetches three 8x18 activation boxes (one per horizontal tap) and,
// when the weights do not fit in smem, the full weight matrix per two tiles.  This kernel
// fetches per tile and channel block
//   * ONE 10(w) x 18(h) pixel box: all nine taps are start-address offsets
//     (dh*10 + dw) * row_bytes into it; a tile row is 8 consecutive pixels = one 8-row swizzle
//     atom that starts on a 128-byte (not 1024-byte) boundary, atoms SBO = 10 rows apart.
//     Verified on B200: the 128B/64B swizzle is a function of the absolute smem address (TMA
//     writes it that way and the UMMA unit reads it that way, descriptor base_offset = 0), so
//     the shifted windows stay canonical.  A traffic per
//     tile drops from 3*18*8 = 432 to 180 pixel rows (1.4x the tile itself).
//   * streamed weights through their own ring of one-tap tiles, shared by MT = 2 pixel tiles and
//     MULTICAST across a thread-block cluster (each CTA fetches 1/CL of every tile), so weight
//     traffic per CTA drops by MT*CL; or resident weights when the packed matrix fits.
//
// Roles: warp 0 = activation TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = weight TMA producer, warps 4-11 = epilogue (tcgen05.ld -> scale/bias -> PixelNorm
// -> LeakyReLU -> bf16 -> swizzled smem -> TMA store).  Accumulators double-buffered in TMEM.
// Replaces cuDNN fprop/dgrad of EqualConv2d 3x3 (reference progan_modules.py:63-73,135,139).
#include "tc_common.cuh"
#include <stdlib.h>

namespace pg {
namespace tc {

struct Conv4Params {
  int N, H, W, Cin, Cout;
  int tiles_w, tiles_h, num_tiles, num_super;
  int ncb;
  int a_stages, w_stages;
  int box_pad;                         // smem bytes per activation box (multiple of 1024)
  int wtile_bytes;                     // one weight tap tile: Cout x BK
  int tmem_cols;
  int epi;
  float scale, slope;
  const float *bias;
  float *r_out;
  // fused activation backward (ABW kernels): the conv output is dh, the gradient w.r.t. the
  // activation y_prev = lrelu(pixelnorm(a_prev)); the epilogue turns it into da_prev
  const __nv_bfloat16 *y_prev;         // [N,H,W,Cout]
  const float *r_prev;                 // [N,H,W] PixelNorm rsqrt (use_pn) or nullptr
  float *colsum;                       // += per-channel sum of da_prev (bias gradient) or nullptr
  int use_pn;
  int pool;                            // also write the 2x2-average-pooled activation (tmap_yp)
  int dbg;                             // experiment knobs (PG_DBG)
};

constexpr int kC4EpiThreads = 256;                 // 8 epilogue warps: 2 per TMEM lane quadrant
constexpr int kC4Threads = 128 + kC4EpiThreads;    // + 4 control warps
constexpr int kC4MaxA = 8, kC4MaxW = 12;

template <int BK, int NCB, int MT, bool RES, int COUT, int CL, bool ABW>
__global__ void __launch_bounds__(kC4Threads, 1)
conv4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_yp,
                const Conv4Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t row_bytes = BK * 2u;                          // 128 or 64
  constexpr uint32_t box_real = 18u * 10u * row_bytes;
  constexpr uint32_t kBoxPad = (box_real + 1023u) / 1024u * 1024u;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_bytes = RES ? (uint32_t)(9 * p.ncb * p.wtile_bytes)
                               : (uint32_t)(p.w_stages * p.wtile_bytes);
  const uint32_t smem_w = base;
  const uint32_t smem_a = base + w_bytes;
  const uint32_t smem_out = smem_a + (uint32_t)(p.a_stages * MT) * kBoxPad;
  const uint32_t out_bytes = 128u * (uint32_t)COUT * 2u;
  // pooled staging tile (32 pooled pixels x COUT) only when the launch asks for it
  const uint32_t smem_pool = smem_out + out_bytes;
  const uint32_t bar_base = smem_pool + (p.pool ? out_bytes / 4u : 0u);
  auto afull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bar_base + 8u * (uint32_t)(kC4MaxA + s); };
  auto wfull = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kC4MaxA + s); };
  auto wempty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kC4MaxA + kC4MaxW + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kC4MaxA + 2 * kC4MaxW + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kC4MaxA + 2 * kC4MaxW + 2 + a); };
  const uint32_t wres_bar = bar_base + 8u * (uint32_t)(2 * kC4MaxA + 2 * kC4MaxW + 4);
  const uint32_t tmem_slot = wres_bar + 8u;
  const uint32_t bias_s = (tmem_slot + 16u + 15u) & ~15u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gbase + (tmem_slot - base));
  float *bias_ptr = reinterpret_cast<float *>(gbase + (bias_s - base));
  float *ss_buf = bias_ptr + 128;                    // [2][128] partial per-pixel reductions
  uint8_t *out_ptr = gbase + (smem_out - base);
  uint8_t *pool_ptr = gbase + (smem_pool - base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int cid = CL > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int ncl = (int)gridDim.x / CL;
  constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_y);
    if (p.pool) prefetch_tmap(&tmap_yp);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < p.w_stages; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), CL);          // every CTA of the cluster must have drained the slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), kC4EpiThreads);
    }
    mbar_init(wres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int c = threadIdx.x; c < COUT; c += kC4Threads) bias_ptr[c] = p.bias ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peers' barriers exist before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr int acc_stride = MT * COUT;       // TMEM columns per accumulator stage

  if (warp == 0) {
    // ===================== activation producer =====================
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      for (int sb = cid * CL; sb < p.num_super; sb += ncl * CL) {
        const int st = sb + (int)rank;
        for (int cb = 0; cb < NCB; ++cb) {
          mbar_wait(aempty(as), aph ^ 1u);
          if (p.dbg & 8) {            // experiment: no activation loads
            mbar_arrive(afull(as));
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            continue;
          }
          mbar_expect_tx(afull(as), (uint32_t)MT * box_real);
          for (int mt = 0; mt < MT; ++mt) {
            const int tile = st * MT + mt;
            const int tw = tile % p.tiles_w;
            const int th = (tile / p.tiles_w) % p.tiles_h;
            const int n = tile / (p.tiles_w * p.tiles_h);
            tma_load_4d(smem_a + (uint32_t)(as * MT + mt) * kBoxPad, &tmap_x, afull(as), cb * BK,
                        tw * 8 - 1, th * 16 - 1, n);
          }
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer =====================
    if (lane == 0) {
      if (RES) {
        mbar_expect_tx(wres_bar, w_bytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int cb = 0; cb < p.ncb; ++cb)
            tma_load_2d(smem_w + (uint32_t)((tap * p.ncb + cb) * p.wtile_bytes), &tmap_w, wres_bar,
                        tap * p.Cin + cb * BK, 0);
      } else {
        int ws = 0;
        uint32_t wph = 0;
        constexpr uint32_t part_bytes = (uint32_t)(COUT / CL) * row_bytes;
        for (int sb = cid * CL; sb < p.num_super; sb += ncl * CL) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(wempty(ws), wph ^ 1u);
              mbar_expect_tx(wfull(ws), (uint32_t)p.wtile_bytes);
              const uint32_t slot = smem_w + (uint32_t)(ws * p.wtile_bytes);
              if (CL == 1)
                tma_load_2d(slot, &tmap_w, wfull(ws), tap * p.Cin + cb * BK, 0);
              else
                tma_load_2d_mc(slot + rank * part_bytes, &tmap_w, wfull(ws), tap * p.Cin + cb * BK,
                               (int)rank * (COUT / CL), cmask);
              if (++ws == p.w_stages) { ws = 0; wph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Measured (profiles/umma_rate.cu): one M128 x N x K16 SS-mode MMA costs max(N/2, 32 + N/4)
    // cycles, so the issue loop must stay far below ~48 cycles per instruction.  All descriptor
    // arithmetic is therefore compile-time offsets from two per-stage bases, and ONE elected
    // thread issues the 9 taps x MT x BK/16 instructions of a channel block back to back (the
    // election, the warp re-convergence and the register->uniform moves happen once per block,
    // not once per tap).
    constexpr uint32_t layout = row_bytes == 128 ? 2u : 4u;
    constexpr uint32_t sbo_a = 10u * row_bytes;                       // next tile row of the box
    constexpr uint32_t sbo_b = 8u * row_bytes;
    constexpr int nk = BK / 16;
    constexpr uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    constexpr uint32_t hi_b = ((sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    constexpr uint32_t lbo_lo = 1u << 16;
    constexpr uint32_t wtile16 = (uint32_t)(COUT * BK * 2) >> 4;
    constexpr uint32_t box16 = kBoxPad >> 4;
    const uint32_t idesc = make_idesc_bf16(128, COUT, 0, 0);
    if (RES) {
      mbar_wait(wres_bar, 0);
      tc_fence_after();
    }
    // The tensor pipe queues only ~2 MMAs (measured: every ~100 cycles the issuing thread spends
    // in a barrier wait between two MMAs shows up as a bubble), so the loop is software-pipelined:
    // ONE thread is elected for the whole kernel, and the barrier waits that the NEXT tap /
    // channel block / tile needs are performed just before the LAST TWO MMAs of the current tap,
    // i.e. while two instructions are still in flight.
    if (elect_one_sync()) {
      int as = 0, ws = 0, acc = 0;
      uint32_t aph = 0, wph = 0, acc_phase = 0;
      const int stride = ncl * CL;
      int sb = cid * CL;
      if (sb < p.num_super) {               // prologue: what the very first MMA needs
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        mbar_wait(afull(as), aph);
        if (!RES) mbar_wait(wfull(ws), wph);
        tc_fence_after();
      }
      for (; sb < p.num_super; sb += stride) {
        const bool has_next_tile = sb + stride < p.num_super;
        const uint32_t d_base = tmem_base + (uint32_t)(acc * acc_stride);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          const uint32_t a16 = ((smem_a + (uint32_t)(as * MT) * kBoxPad) >> 4) | lbo_lo;
          const uint32_t w16 = (smem_w >> 4) | lbo_lo;
          const int as_next = (as + 1 == p.a_stages) ? 0 : as + 1;
          const uint32_t aph_next = (as + 1 == p.a_stages) ? (aph ^ 1u) : aph;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dh = tap / 3, dw = tap % 3;
            const uint32_t b16 = RES ? w16 + (uint32_t)(tap * NCB + cb) * wtile16
                                     : w16 + (uint32_t)ws * wtile16;
            const uint32_t a_off16 = ((uint32_t)(dh * 10 + dw) * row_bytes) >> 4;
            constexpr int NM = MT * nk;                  // MMAs of one tap
            constexpr int SPLIT = NM > 2 ? NM - 2 : 0;   // the waits go in front of the last two
            const int ws_next = (ws + 1 == p.w_stages) ? 0 : ws + 1;
            const uint32_t wph_next = (ws + 1 == p.w_stages) ? (wph ^ 1u) : wph;
            // barriers the next tap / block / tile needs: polled ONCE in front of the last two
            // MMAs (hidden behind the in-flight ones when they are already complete, the common
            // case when MMA-bound), waited for after this tap's commits otherwise — a blocking
            // wait in front of the last MMAs would delay this tile's own completion
            const bool last_of_tile = (tap == 8) && (cb == NCB - 1);
            const bool need_w = !RES && !(last_of_tile && !has_next_tile);
            const bool need_a = (tap == 8) && (cb < NCB - 1 || has_next_tile);
            const bool need_t = last_of_tile && has_next_tile;
            const uint32_t t_par = acc ? (acc_phase ^ 1u) ^ 1u : acc_phase ^ 1u;
            bool ok_w = !need_w, ok_a = !need_a, ok_t = !need_t;
#pragma unroll
            for (int j = 0; j < NM; ++j) {
              if (j == SPLIT && (!RES || tap == 8)) {
                if (need_w) ok_w = mbar_try_wait(wfull(ws_next), wph_next);
                if (need_a) ok_a = mbar_try_wait(afull(as_next), aph_next);
                if (need_t) ok_t = mbar_try_wait(tempty(acc ^ 1), t_par);
                tc_fence_after();
              }
              const int mt = j / nk, k = j % nk;
              if (!(p.dbg & 16)) {       // experiment: no MMAs
                const uint64_t ad = ((uint64_t)hi_a << 32) |
                                    (uint64_t)(a16 + (uint32_t)mt * box16 + a_off16 + (uint32_t)(k * 2));
                const uint64_t bd = ((uint64_t)hi_b << 32) | (uint64_t)(b16 + (uint32_t)(k * 2));
                umma_bf16(d_base + (uint32_t)(mt * COUT), ad, bd, idesc, (cb | tap | k) ? 1u : 0u);
              }
            }
            if (!RES) {
              if (CL == 1) umma_commit(wempty(ws));
              else umma_commit_mc(wempty(ws), cmask);
              ws = ws_next;
              wph = wph_next;
            }
            if (tap == 8) {
              umma_commit(aempty(as));
              if (cb == NCB - 1) umma_commit(tfull(acc));
            }
            if (!(ok_w && ok_a && ok_t)) {
              if (!ok_w) mbar_wait(wfull(ws), wph);            // ws / wph already advanced
              if (!ok_a) mbar_wait(afull(as_next), aph_next);
              if (!ok_t) mbar_wait(tempty(acc ^ 1), t_par);
              tc_fence_after();
            }
          }
          as = as_next;
          aph = aph_next;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // Two warps per TMEM lane quadrant, each owning half of the channel columns of its 32
    // pixels: one tcgen05.ld pass, values stay in registers; per-pixel reductions over the
    // channels (PixelNorm sum of squares, <p,u> of the fused backward) are completed through a
    // 1 KB smem exchange between the two halves.  (16 epilogue warps were measured slower.)
    constexpr int CPT = COUT / 2;                 // columns per thread: 16 / 32 / 64
    constexpr int out_chunk = (COUT % 64 == 0) ? 64 : 32;
    constexpr int chunk_rows_bytes = out_chunk * 2;
    constexpr int swz_bits = chunk_rows_bytes == 128 ? 3 : 2;
    constexpr int n_chunks = COUT / out_chunk;
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;             // 0..1
    const int row = q * 32 + lane;                // tile row: pixel (hl = row/8, wl = row%8)
    const int et = threadIdx.x - 128;             // 0..255
    const int col0 = part * CPT;
    const float invC = 1.f / (float)COUT;
    const float scale = p.scale, slope = p.slope;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float inv_slope = 1.f / slope;
    float csum[2] = {0.f, 0.f};                   // ABW: this lane's share of the bias gradient
    // ABW: this thread's slice of y_prev (and r_prev) for a tile is loaded one tile ahead, right
    // after the previous tile's arithmetic, and prefetched into L2 two tiles ahead: the HBM
    // latency hides behind the accumulator wait instead of sitting on the epilogue's critical path
    uint4 yraw[ABW ? CPT / 8 : 1];
    float rp = 1.f;
    auto pix_of = [&](int tile) -> long long {
      const int tw_ = tile % p.tiles_w;
      const int th_ = (tile / p.tiles_w) % p.tiles_h;
      const int n_ = tile / (p.tiles_w * p.tiles_h);
      return ((long long)n_ * p.H + (th_ * 16 + (row >> 3))) * p.W + tw_ * 8 + (row & 7);
    };
    auto next_tile = [&](int sb_, int mt_, int ahead) -> int {     // flat successor, -1 past the end
      for (int a = 0; a < ahead; ++a) {
        if (++mt_ == MT) { mt_ = 0; sb_ += ncl * CL; }
      }
      return sb_ < p.num_super ? (sb_ + (int)rank) * MT + mt_ : -1;
    };
    auto load_y = [&](int tile) {
      if (tile < 0) return;
      const long long px = pix_of(tile);
      const uint4 *yp = reinterpret_cast<const uint4 *>(p.y_prev + px * COUT + col0);
#pragma unroll
      for (int i = 0; i < CPT / 8; ++i) yraw[i] = __ldg(yp + i);
      if (p.use_pn) rp = __ldg(p.r_prev + px);
    };
    auto prefetch_y = [&](int tile) {
      if (tile < 0) return;
      const char *yp = reinterpret_cast<const char *>(p.y_prev + pix_of(tile) * COUT + col0);
#pragma unroll
      for (int b = 0; b < CPT * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(yp + b));
    };
    if (ABW) {
      load_y(next_tile(cid * CL, 0, 0));
      prefetch_y(next_tile(cid * CL, 0, 1));
    }
    for (int sb = cid * CL; sb < p.num_super; sb += ncl * CL) {
      const int st = sb + (int)rank;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int tile = st * MT + mt;
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        const long long pix = ((long long)n * p.H + (h0 + (row >> 3))) * p.W + w0 + (row & 7);
        if (ABW) prefetch_y(next_tile(sb, mt, 2));
        if (mt == 0) {
          mbar_wait(tfull(acc), acc_phase);
          tc_fence_after();
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                (uint32_t)(acc * acc_stride + mt * COUT + col0);
        uint32_t vr[CPT];
        float v[CPT];
        float r = 1.f;
        if (!ABW) {
          // accumulator drain in 16-column pieces: the tcgen05.ld of piece c+1 is in flight while
          // piece c is scaled and squared (TMEM reads are 64 B/clk per SM: the drain of a
          // 128x128 fp32 tile alone is ~1000 cycles)
          constexpr int CH = 16, NCHK = CPT / CH;
          float ss = 0.f;
          tmem_ld<CH>(t_addr, vr);
#pragma unroll
          for (int c = 0; c < NCHK; ++c) {
            tmem_ld_wait();
            if (c + 1 < NCHK) {
              tmem_ld<CH>(t_addr + (uint32_t)((c + 1) * CH), vr + (c + 1) * CH);
            } else if (mt == MT - 1) {   // accumulator stage fully read: hand it back to the MMA warp
              tc_fence_before();
              mbar_arrive(tempty(acc));
            }
#pragma unroll
            for (int j = c * CH; j < (c + 1) * CH; j += 4) {
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
          }
          if (p.dbg & 2) continue;       // experiment: accumulator drain only
          if (p.epi == PG_EPI_PN_LRELU) {
            ss_buf[part * 128 + row] = ss;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            r = rsqrtf((ss_buf[row] + ss_buf[128 + row]) * invC + 1e-8f);
          }
        } else {
          tmem_ld<CPT>(t_addr, vr);
          tmem_ld_wait();
          if (mt == MT - 1) {            // accumulator stage fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(tempty(acc));
          }
          // da = r (u - p <p,u>/C), u = m * dh, (p, m) rebuilt from the stored activation
          float s_pu = 0.f;
#pragma unroll
          for (int i = 0; i < CPT / 8; ++i) {
            const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yraw[i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 yy = __bfloat1622float2(h2[e]);
              const int j = i * 8 + 2 * e;
              const float u0 = __uint_as_float(vr[j]) * scale * (yy.x > 0.f ? 1.f : slope);
              const float u1 = __uint_as_float(vr[j + 1]) * scale * (yy.y > 0.f ? 1.f : slope);
              s_pu = fmaf(yy.x > 0.f ? yy.x : yy.x * inv_slope, u0, s_pu);
              s_pu = fmaf(yy.y > 0.f ? yy.y : yy.y * inv_slope, u1, s_pu);
              v[j] = u0;
              v[j + 1] = u1;
            }
          }
          if (p.use_pn) {
            ss_buf[part * 128 + row] = s_pu;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const float k = (ss_buf[row] + ss_buf[128 + row]) * invC;
#pragma unroll
            for (int i = 0; i < CPT / 8; ++i) {
              const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yraw[i]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 yy = __bfloat1622float2(h2[e]);
                const int j = i * 8 + 2 * e;
                v[j] = rp * fmaf(-(yy.x > 0.f ? yy.x : yy.x * inv_slope), k, v[j]);
                v[j + 1] = rp * fmaf(-(yy.y > 0.f ? yy.y : yy.y * inv_slope), k, v[j + 1]);
              }
            }
          }
          load_y(next_tile(sb, mt, 1));          // yraw / rp are free again: fetch the next tile's
        }
        if (et == 0) tma_store_wait_read0();       // staging buffer free again?
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          const int chunk_i = col0 / out_chunk;
          const int cin_chunk = col0 - chunk_i * out_chunk;
          uint8_t *tile_base = out_ptr + (size_t)chunk_i * 128 * chunk_rows_bytes;
#pragma unroll
          for (int i = 0; i < CPT / 8; ++i) {                      // 8 channels = 16 bytes per store
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a0 = v[i * 8 + 2 * e] * r, a1 = v[i * 8 + 2 * e + 1] * r;
              if (!ABW && p.epi != PG_EPI_LINEAR) {
                a0 = a0 > 0.f ? a0 : a0 * slope;
                a1 = a1 > 0.f ? a1 : a1 * slope;
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              pk[e] = *reinterpret_cast<uint32_t *>(&h);
            }
            const uint32_t off = (uint32_t)row * (uint32_t)chunk_rows_bytes +
                                 (uint32_t)cin_chunk * 2u + (uint32_t)i * 16u;
            *reinterpret_cast<uint4 *>(tile_base + swz(off, swz_bits)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        if (!ABW && p.epi == PG_EPI_PN_LRELU && part == 0) p.r_out[pix] = r;
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && !(p.dbg & 1)) {
#pragma unroll
          for (int ch = 0; ch < n_chunks; ++ch)
            tma_store_4d(&tmap_y, smem_out + (uint32_t)ch * 128u * (uint32_t)chunk_rows_bytes,
                         ch * out_chunk, w0, h0, n);
          tma_store_commit();
        }
        if (!ABW && p.pool) {
          // 2x2 average pool of the staged tile (the reference's bilinear x0.5,
          // progan_modules.py:299): 32 pooled pixels x COUT/8 16-byte chunks over 256 threads
          constexpr int CH8 = COUT / 8;
          for (int item = et; item < 32 * CH8; item += kC4EpiThreads) {
            const int pp = item / CH8, c8 = item - pp * CH8;
            const int ph = pp >> 2, pw = pp & 3;
            const int chunk_i = (c8 * 8) / out_chunk;
            const uint32_t cb = (uint32_t)((c8 * 8) % out_chunk) * 2u;
            const uint8_t *tb = out_ptr + (size_t)chunk_i * 128 * chunk_rows_bytes;
            float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int srow = (2 * ph + (q4 >> 1)) * 8 + 2 * pw + (q4 & 1);
              const uint4 raw = *reinterpret_cast<const uint4 *>(
                  tb + swz((uint32_t)srow * (uint32_t)chunk_rows_bytes + cb, swz_bits));
              const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h2[e]);
                acc8[2 * e] += f.x;
                acc8[2 * e + 1] += f.y;
              }
            }
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 h = __floats2bfloat162_rn(0.25f * acc8[2 * e], 0.25f * acc8[2 * e + 1]);
              pk[e] = *reinterpret_cast<uint32_t *>(&h);
            }
            uint8_t *pb = pool_ptr + (size_t)chunk_i * 32 * chunk_rows_bytes;
            *reinterpret_cast<uint4 *>(pb + swz((uint32_t)pp * (uint32_t)chunk_rows_bytes + cb, swz_bits)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0 && !(p.dbg & 1)) {
#pragma unroll
            for (int ch = 0; ch < n_chunks; ++ch)
              tma_store_4d(&tmap_yp, smem_pool + (uint32_t)ch * 32u * (uint32_t)chunk_rows_bytes,
                           ch * out_chunk, w0 >> 1, h0 >> 1, n);
            tma_store_commit();
          }
        }
        if (ABW && p.colsum != nullptr) {
          // per-channel sum over the warp's 32 pixels by a halving butterfly: each step trades
          // half of the remaining channels with the partner lane (CPT - 1 shuffles in total)
          int nrem = CPT;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            if (nrem >= 2) {
              const int hn = nrem / 2;
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < CPT / 2; ++j) {
                if (j < hn) {
                  const float send = up ? v[j] : v[j + hn];
                  const float keep = up ? v[j + hn] : v[j];
                  v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              nrem = hn;
            } else {
              v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
            }
          }
          csum[0] += v[0];
          if (CPT >= 64) csum[1] += v[1];
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (ABW && p.colsum != nullptr) {
      // channel owned by this lane after the butterfly (bit b of the lane picked the upper half
      // at the step with offset 2^b)
      int ch = col0;
      int hn = CPT / 2;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        if (hn >= 1) {
          if (lane & off) ch += hn;
          hn >>= 1;
        }
      }
      if (CPT >= 64) {
        atomicAdd(p.colsum + ch, csum[0]);
        atomicAdd(p.colsum + ch + 1, csum[1]);
      } else if (CPT == 32) {
        atomicAdd(p.colsum + ch, csum[0]);
      } else if ((lane & 1) == 0) {         // CPT = 16: the last step was a plain add
        atomicAdd(p.colsum + ch, csum[0]);
      }
    }
    if (et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA exits while a peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int BK, int NCB, int MT, bool RES, int COUT, int CL, bool ABW>
static cudaError_t launch_c4(const CUtensorMap &tx, const CUtensorMap &tw, const CUtensorMap &ty,
                             const CUtensorMap &typ, const Conv4Params &p, size_t smem,
                             cudaStream_t stream) {
  auto kern = conv4_tc_kernel<BK, NCB, MT, RES, COUT, CL, ABW>;
  static bool attr_set = false;
  static int max_ctas = 0;
  cudaError_t e;
  if (!attr_set) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kC4Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  if (max_ctas == 0) {
    max_ctas = sm_count();
    if (CL > 1) {
      int ncl = 0;
      cfg.gridDim = dim3((unsigned)(sm_count() / CL * CL));
      if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) == cudaSuccess && ncl > 0)
        max_ctas = ncl * CL;
      else
        (void)cudaGetLastError();
      if (max_ctas > sm_count()) max_ctas = sm_count() / CL * CL;
    }
  }
  int grid = p.num_super < max_ctas ? p.num_super : max_ctas;
  grid = grid / CL * CL;
  cfg.gridDim = dim3((unsigned)grid);
  return cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, typ, p);
}

}  // namespace tc

// Returns PG_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible,
// so pg_conv_tc can fall through to the older kernels.
int conv4_tc_launch(const void *x, const void *wp, const float *bias, void *y, float *r_out, int N,
                    int H, int W, int Cin, int Cout, float scale, int epi, float slope,
                    cudaStream_t stream, const void *y_prev, const float *r_prev, float *colsum,
                    int use_pn, void *y_pool) {
  const bool abw = y_prev != nullptr;
  if (abw && y_pool) return PG_ERR_UNSUPPORTED;
  if (const char *e = getenv("PG_CONV_V4"))
    if (atoi(e) == 0) return PG_ERR_UNSUPPORTED;
  int min_h = 16;
  if (const char *e = getenv("PG_C4_MINH")) min_h = atoi(e);
  if (H % 16 || H < min_h || W % 8 || !(Cout == 32 || Cout == 64 || Cout == 128) || Cin % 32)
    return PG_ERR_UNSUPPORTED;
  tc::Conv4Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * N;
  const int BK = (Cin % 64 == 0) ? 64 : 32;
  p.ncb = Cin / BK;
  const int box_real = 18 * 10 * BK * 2;
  p.box_pad = (box_real + 1023) / 1024 * 1024;
  p.wtile_bytes = Cout * BK * 2;
  p.epi = epi; p.scale = scale; p.slope = slope; p.bias = bias; p.r_out = r_out;
  p.y_prev = (const __nv_bfloat16 *)y_prev; p.r_prev = r_prev; p.colsum = colsum; p.use_pn = use_pn;
  p.pool = y_pool != nullptr;
  p.dbg = 0;
  if (const char *e = getenv("PG_DBG")) p.dbg = atoi(e);
  const int out_bytes = 128 * Cout * 2 + (y_pool ? 32 * Cout * 2 : 0);   // staging (+ pooled tile)
  const int misc = 1024 + 8 * (2 * tc::kC4MaxA + 2 * tc::kC4MaxW + 5) + 16 + 16 + 128 * 4 + 2 * 128 * 4 + 64;
  const int budget = 227 * 1024 - out_bytes - misc;
  const int wres = 9 * p.ncb * p.wtile_bytes;
  int force_res = -1, force_mt = -1, force_cl = -1;
  if (const char *e = getenv("PG_C4_RES")) force_res = atoi(e);
  if (const char *e = getenv("PG_C4_MT")) force_mt = atoi(e);
  if (const char *e = getenv("PG_C4_CL")) force_cl = atoi(e);
  int res = (wres + 2 * p.box_pad <= budget) ? 1 : 0;
  if (force_res >= 0 && (force_res == 0 || wres + 2 * p.box_pad <= budget)) res = force_res;
  if (BK == 32 && !res) return PG_ERR_UNSUPPORTED;
  if (!res && Cout < 64) return PG_ERR_UNSUPPORTED;
  int MT, CL;
  if (res) {
    CL = 1;
    // two pixel tiles per accumulator stage when they fit: per-tile handshakes amortise
    MT = (p.num_tiles % 2 == 0 && 4 * Cout <= 512 && wres + 4 * p.box_pad <= budget) ? 2 : 1;
    if (force_mt == 1) MT = 1;
    p.a_stages = (budget - wres) / (MT * p.box_pad);
    if (p.a_stages > tc::kC4MaxA) p.a_stages = tc::kC4MaxA;
    p.w_stages = 0;
  } else {
    MT = (p.num_tiles % 2 == 0 && p.num_tiles / 2 >= sm_count() && 4 * Cout <= 512) ? 2 : 1;
    if (force_mt > 0 && p.num_tiles % force_mt == 0 && 2 * force_mt * Cout <= 512) MT = force_mt;
    const int ns = p.num_tiles / MT;
    CL = (MT == 1 && ns % 4 == 0) ? 4 : (ns % 2 == 0 ? 2 : 1);
    if (force_cl > 0 && ns % force_cl == 0) CL = force_cl;
    p.a_stages = (MT == 1) ? 3 : 2;
    p.w_stages = (budget - p.a_stages * MT * p.box_pad) / p.wtile_bytes;
    if (p.w_stages > tc::kC4MaxW) p.w_stages = tc::kC4MaxW;
    if (p.w_stages < 3) return PG_ERR_UNSUPPORTED;
  }
  if (p.a_stages < 2) return PG_ERR_UNSUPPORTED;
  p.num_super = p.num_tiles / MT;
  int cols = 2 * MT * Cout;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const size_t smem = (size_t)(res ? wres : p.w_stages * p.wtile_bytes) +
                      (size_t)p.a_stages * MT * p.box_pad + out_bytes + misc;

  CUtensorMap tx, tw_, ty;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)BK, 10u, 18u, 1u};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, BK * 2, "pg_conv_tc/v4(x)")) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(Cout / CL)};
    if (int rc = make_tmap_bf16(&tw_, wp, 2, dims, str, box, BK * 2, "pg_conv_tc/v4(w)")) return rc;
  }
  const int out_chunk = (Cout % 64 == 0) ? 64 : 32;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {(uint32_t)out_chunk, 8u, 16u, 1u};
    if (int rc = make_tmap_bf16(&ty, y, 4, dims, str, box, out_chunk * 2, "pg_conv_tc/v4(y)")) return rc;
  }
  CUtensorMap typ = ty;
  if (y_pool) {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)(W / 2) * Cout * 2, (uint64_t)(H / 2) * (W / 2) * Cout * 2};
    uint32_t box[4] = {(uint32_t)out_chunk, 4u, 8u, 1u};
    if (int rc = make_tmap_bf16(&typ, y_pool, 4, dims, str, box, out_chunk * 2, "pg_conv_tc/v4(y_pool)")) return rc;
  }
  cudaError_t e = cudaErrorInvalidValue;
  bool matched = false;
#define PG_C4_TRY(BK_, NCB_, MT_, RES_, CO_, CL_)                                                \
  if (!matched && BK == BK_ && p.ncb == NCB_ && MT == MT_ && res == (RES_ ? 1 : 0) && Cout == CO_ && \
      CL == CL_) {                                                                               \
    matched = true;                                                                              \
    e = abw ? tc::launch_c4<BK_, NCB_, MT_, RES_, CO_, CL_, true>(tx, tw_, ty, typ, p, smem, stream)  \
            : tc::launch_c4<BK_, NCB_, MT_, RES_, CO_, CL_, false>(tx, tw_, ty, typ, p, smem, stream); \
  }
#define PG_C4_RESIDENT(BK_, NCB_, CO_)                                                           \
  PG_C4_TRY(BK_, NCB_, 1, true, CO_, 1) PG_C4_TRY(BK_, NCB_, 2, true, CO_, 1)
#define PG_C4_STREAM(NCB_, CO_)                                                                  \
  PG_C4_TRY(64, NCB_, 1, false, CO_, 1) PG_C4_TRY(64, NCB_, 1, false, CO_, 2)                    \
  PG_C4_TRY(64, NCB_, 1, false, CO_, 4) PG_C4_TRY(64, NCB_, 2, false, CO_, 1)                    \
  PG_C4_TRY(64, NCB_, 2, false, CO_, 2) PG_C4_TRY(64, NCB_, 2, false, CO_, 4)
  PG_C4_RESIDENT(64, 1, 32) PG_C4_RESIDENT(64, 1, 64) PG_C4_RESIDENT(64, 1, 128)
  PG_C4_RESIDENT(64, 2, 32) PG_C4_RESIDENT(64, 2, 64)
  PG_C4_RESIDENT(32, 1, 32) PG_C4_RESIDENT(32, 1, 64) PG_C4_RESIDENT(32, 1, 128)
  PG_C4_STREAM(2, 64) PG_C4_STREAM(2, 128) PG_C4_STREAM(1, 128)
#undef PG_C4_STREAM
#undef PG_C4_RESIDENT
#undef PG_C4_TRY
  if (!matched) return PG_ERR_UNSUPPORTED;
  if (e != cudaSuccess) {
    set_error("pg_conv_tc/v4: launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pg_conv_tc/v4: CUDA launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace pg
