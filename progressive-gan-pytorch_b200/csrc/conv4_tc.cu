// 3x3 pad-1 implicit-GEMM convolution, third generation ("single halo box") kernel.
//
// A kernel that fetches one activation box per horizontal tap (three 8x18 boxes per 128-pixel
// tile and 64-channel block) and the whole weight matrix per two tiles is bound by L2->SM
// bandwidth (~30-40 B/cycle/SM measured in round 1).  This kernel fetches per tile and channel
// block
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

constexpr int kC4EpiThreads = 256;                 // 8 epilogue warps: two groups of 4 (one per TMEM stage)
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
  // PAIR: the two CTAs of a cluster execute ONE M = 256 MMA (tcgen05.mma.cta_group::2): each
  // CTA feeds its own 128-pixel tile (A) and HALF of the weight rows (B).  Per MMA a CTA then
  // reads 4 KB of A + N*16 B of B instead of 4 KB + N*32 B (smem operand bandwidth, 128 B/clk, is
  // what bounds every N <= 128 layer), and half of the packed weights per CTA is small enough to
  // stay RESIDENT for every layer of the network (128 -> 128: 144 KB), which removes the weight
  // stream — the other large consumer of smem bandwidth — altogether.  Rank 0 (leader) issues
  // all MMAs; both CTAs run an activation producer and the epilogue for their own tile.
  //   afull[s]  (leader)  count 2: each CTA's producer arms it with its box bytes; both TMA
  //                       loads (.cta_group::2) complete_tx on the leader's barrier
  //   aempty[s], tfull[a] (each CTA) count 1: tcgen05.commit.cta_group::2, multicast to the pair
  //   tempty[a] (leader)  count 8: one arrive per epilogue warp of group a in both CTAs
  constexpr bool PAIR = RES && CL == 2;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_bytes = RES ? (uint32_t)(9 * p.ncb * p.wtile_bytes) / (PAIR ? 2u : 1u)
                               : (uint32_t)(p.w_stages * p.wtile_bytes);
  const uint32_t smem_w = base;
  const uint32_t smem_a = base + w_bytes;
  const uint32_t smem_out = smem_a + (uint32_t)(p.a_stages * MT) * kBoxPad;
  // one staging tile (128 pixels x min(COUT, 64) channels) per epilogue group, and one pooled
  // staging tile (32 pooled pixels) per group when the launch asks for the fused average pool
  // (two of each per group for COUT <= 64: a TMA store needs ~1000 cycles to read its tile, longer
  // than the whole epilogue of a 128 x 32 tile, so with ONE tile per group the next tile's staging
  // writes waited for it every time - measured 920 cycles per 128 x 32 tile for the epilogue alone)
  constexpr uint32_t kHC = COUT > 64 ? 64u : (uint32_t)COUT;
  constexpr uint32_t kNSB = COUT > 64 ? 1u : 2u;
  constexpr uint32_t out_bytes = 2u * kNSB * 128u * kHC * 2u;
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
      mbar_init(afull(s), PAIR ? 2 : 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < p.w_stages; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), CL);          // every CTA of the cluster must have drained the slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), PAIR ? 8 : 4);          // one arrive per warp of the group that owns stage a
    }
    mbar_init(wres_bar, PAIR ? 2 : 1);
    fence_barrier_init();
  }
  if (!PAIR && warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // everything above touches only this CTA's shared memory / TMEM: it overlaps the tail of the
  // previous kernel; from here on the predecessors' results are read
  pg::grid_dep_sync();
  for (int c = threadIdx.x; c < COUT; c += kC4Threads) bias_ptr[c] = p.bias ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peers' barriers exist before any multicast lands
  if (PAIR) {                              // pair allocation: both CTAs, after the cluster is up
    if (warp == 2) tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const bool leader = rank == 0;
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
            if (!PAIR || leader) mbar_arrive(afull(as));
            else mbar_arrive_remote(mapa_shared(afull(as), 0));
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            continue;
          }
          if (!PAIR || leader) mbar_expect_tx(afull(as), (uint32_t)MT * box_real);
          else mbar_expect_tx_remote(mapa_shared(afull(as), 0), (uint32_t)MT * box_real);
          for (int mt = 0; mt < MT; ++mt) {
            const int tile = st * MT + mt;
            const int tw = tile % p.tiles_w;
            const int th = (tile / p.tiles_w) % p.tiles_h;
            const int n = tile / (p.tiles_w * p.tiles_h);
            const uint32_t dst = smem_a + (uint32_t)(as * MT + mt) * kBoxPad;
            if (PAIR) tma_load_4d_pair(dst, &tmap_x, afull(as), cb * BK, tw * 8 - 1, th * 16 - 1, n);
            else tma_load_4d(dst, &tmap_x, afull(as), cb * BK, tw * 8 - 1, th * 16 - 1, n);
          }
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer =====================
    if (lane == 0) {
      if (PAIR) {                 // this CTA's half of the rows of every tap tile, once
        const uint32_t whalf = (uint32_t)p.wtile_bytes / 2u;
        if (leader) mbar_expect_tx(wres_bar, w_bytes);
        else mbar_expect_tx_remote(mapa_shared(wres_bar, 0), w_bytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int cb = 0; cb < p.ncb; ++cb)
            tma_load_2d_pair(smem_w + (uint32_t)(tap * p.ncb + cb) * whalf, &tmap_w, wres_bar,
                             tap * p.Cin + cb * BK, (int)rank * (COUT / 2));
      } else if (RES) {
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
    constexpr uint32_t wtile16 = (uint32_t)((PAIR ? COUT / 2 : COUT) * BK * 2) >> 4;
    constexpr uint32_t box16 = kBoxPad >> 4;
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, COUT, 0, 0);
    if (RES && (!PAIR || leader)) {
      mbar_wait(wres_bar, 0);
      tc_fence_after();
    }
    // The tensor pipe queues only ~2 MMAs (measured: every ~100 cycles the issuing thread spends
    // in a barrier wait between two MMAs shows up as a bubble), so the loop is software-pipelined:
    // ONE thread is elected for the whole kernel, and the barrier waits that the NEXT tap /
    // channel block / tile needs are performed just before the LAST TWO MMAs of the current tap,
    // i.e. while two instructions are still in flight.
    if ((!PAIR || leader) && elect_one_sync()) {
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
                if (PAIR) umma_bf16_2cta(d_base + (uint32_t)(mt * COUT), ad, bd, idesc, (cb | tap | k) ? 1u : 0u);
                else umma_bf16(d_base + (uint32_t)(mt * COUT), ad, bd, idesc, (cb | tap | k) ? 1u : 0u);
              }
            }
            if (!RES) {
              if (CL == 1) umma_commit(wempty(ws));
              else umma_commit_mc(wempty(ws), cmask);
              ws = ws_next;
              wph = wph_next;
            }
            if (tap == 8) {
              if (PAIR) {
                umma_commit2(aempty(as));
                if (cb == NCB - 1) umma_commit2(tfull(acc));
              } else {
                umma_commit(aempty(as));
                if (cb == NCB - 1) umma_commit(tfull(acc));
              }
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
    // ===================== epilogue: two independent groups of 4 warps =====================
    // Group g owns accumulator stage g (the MMA warp alternates stages per super tile), so the
    // two groups work on different tiles at the same time and never synchronise with each
    // other.  Inside a group, warp q reads TMEM lane quadrant q and every thread owns ONE pixel
    // with its WHOLE channel vector: the PixelNorm sum of squares (and <p,u> of the fused
    // backward) is a per-thread reduction - no cross-warp exchange.  Each group has its own
    // staging tile and issues its own TMA stores; the only synchronisation per tile is two
    // 128-thread named barriers around the staging writes.  (The first version of this
    // epilogue split a row's channels over two warps, exchanged partial sums through smem and
    // shared ONE staging tile between all 8 warps: 1700 cycles per 128x32 tile, 3800-5000 per
    // 128x128 tile - profiles/r2/diag_conv4_before.txt - which bound every layer with
    // Cout <= 64.)  128 output channels are processed as two 64-column passes over TMEM (the
    // second pass re-reads its columns; a TMEM read costs far less than the registers would).
    constexpr int HC = COUT > 64 ? 64 : COUT;       // columns per pass = channels per staging tile
    constexpr int NH = COUT / HC;                   // 1, or 2 for COUT = 128
    constexpr int row_b = HC * 2;                   // staging row bytes: 64 or 128
    constexpr int swz_bits = row_b == 128 ? 3 : 2;
    const int g = (warp - 4) >> 2;                  // group = accumulator stage
    const int q = warp & 3;
    const int row = q * 32 + lane;                  // tile row: pixel (hl = row/8, wl = row%8)
    const int gt = (int)threadIdx.x - 128 - g * 128;   // 0..127 inside the group
    const int bar_id = 1 + g;
    constexpr int NSB = (int)kNSB;                  // staging tiles per group (alternating)
    int sbuf = 0;
    const float invC = 1.f / (float)COUT;
    const float scale = p.scale, slope = p.slope;
    const float inv_slope = 1.f / slope;
    const bool pn = ABW ? (p.use_pn != 0) : (p.epi == PG_EPI_PN_LRELU);
    const bool act = !ABW && p.epi != PG_EPI_LINEAR;
    auto group_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    // this warp has read its quadrant of accumulator stage g: one arrive per warp on the MMA
    // issuer's barrier (the pair's leader owns it)
    auto release_stage = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || leader) mbar_arrive(tempty(g));
        else mbar_arrive_remote(mapa_shared(tempty(g), 0));
      }
    };
    float csum[2 * NH];                             // ABW: this lane's share of the bias gradient
#pragma unroll
    for (int i = 0; i < 2 * NH; ++i) csum[i] = 0.f;
    uint32_t ph = 0;
    int it = 0;
    for (int sb = cid * CL; sb < p.num_super; sb += ncl * CL, ++it) {
      if ((it & 1) != g) continue;
      const int st = sb + (int)rank;
      mbar_wait(tfull(g), ph);
      ph ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int tile = st * MT + mt;
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        const long long pix = ((long long)n * p.H + (h0 + (row >> 3))) * p.W + w0 + (row & 7);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                (uint32_t)(g * acc_stride + mt * COUT);
        uint8_t *stage_ptr = out_ptr + (size_t)(g * NSB + sbuf) * 128 * row_b;
        const uint32_t stage_s = smem_out + (uint32_t)(g * NSB + sbuf) * 128u * (uint32_t)row_b;
        uint8_t *pool_g = pool_ptr + (size_t)(g * NSB + sbuf) * 32 * row_b;
        const uint32_t pool_s = smem_pool + (uint32_t)(g * NSB + sbuf) * 32u * (uint32_t)row_b;
        if (NSB > 1) sbuf ^= 1;
        float red = 0.f;                     // sum of squares (forward) or <p,u> (fused backward)
        if constexpr (ABW && NH == 2) {
          // fused backward, 128 channels: columns HC..COUT-1 contribute to <p,u> only in this pass
          // (16 columns and the matching 32 bytes of the stored activation at a time; done first so
          // that these temporaries are dead when the first half is held in registers)
          const uint4 *yp2 = reinterpret_cast<const uint4 *>(p.y_prev + pix * COUT + HC);
          uint32_t tb[2][16];
          uint4 yb[2][2];
          tmem_ld<16>(t_addr + (uint32_t)HC, tb[0]);
          yb[0][0] = __ldg(yp2);
          yb[0][1] = __ldg(yp2 + 1);
#pragma unroll
          for (int c = 0; c < HC / 16; ++c) {
            tmem_ld_wait();
            if (c + 1 < HC / 16) {
              tmem_ld<16>(t_addr + (uint32_t)(HC + (c + 1) * 16), tb[(c + 1) & 1]);
              yb[(c + 1) & 1][0] = __ldg(yp2 + 2 * (c + 1));
              yb[(c + 1) & 1][1] = __ldg(yp2 + 2 * (c + 1) + 1);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yb[c & 1][i]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 yy = __bfloat1622float2(h2[e]);
                const int j = i * 8 + 2 * e;
                const float u0 = __uint_as_float(tb[c & 1][j]) * scale * (yy.x > 0.f ? 1.f : slope);
                const float u1 = __uint_as_float(tb[c & 1][j + 1]) * scale * (yy.y > 0.f ? 1.f : slope);
                red = fmaf(yy.x > 0.f ? yy.x : yy.x * inv_slope, u0, red);
                red = fmaf(yy.y > 0.f ? yy.y : yy.y * inv_slope, u1, red);
              }
            }
          }
        }
        float v[HC];                         // accumulators are loaded in place (one register set)
        uint32_t *vr = reinterpret_cast<uint32_t *>(v);
        // ---- pass 1: the first HC columns stay in registers
#pragma unroll
        for (int c = 0; c < HC; c += 32) tmem_ld<32>(t_addr + (uint32_t)c, vr + c);
        uint4 yraw[ABW ? HC / 8 : 1];
        float rp = 1.f;
        if constexpr (ABW) {   // stored activation of the layer in front (this thread's pixel)
          const uint4 *yp = reinterpret_cast<const uint4 *>(p.y_prev + pix * COUT);
#pragma unroll
          for (int i = 0; i < HC / 8; ++i) yraw[i] = __ldg(yp + i);
          if (pn) rp = __ldg(p.r_prev + pix);
        }
        // this staging tile is free once the TMA store(s) that last used it have read it: with
        // two tiles per group the previous tile's stores may still be in flight
        if (gt == 0) {
          if (NSB == 1) tma_store_wait_read0();
          else if (p.pool) tma_store_wait_read<2>();
          else tma_store_wait_read<1>();
        }
        group_bar();
        tmem_ld_wait();
        if (NH == 1 && mt == MT - 1) release_stage();   // stage fully read: back to the MMA warp
        if constexpr (!ABW) {
#pragma unroll
          for (int j = 0; j < HC; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4 *>(bias_ptr + j);
            v[j] = fmaf(__uint_as_float(vr[j]), scale, b4.x);
            v[j + 1] = fmaf(__uint_as_float(vr[j + 1]), scale, b4.y);
            v[j + 2] = fmaf(__uint_as_float(vr[j + 2]), scale, b4.z);
            v[j + 3] = fmaf(__uint_as_float(vr[j + 3]), scale, b4.w);
            red = fmaf(v[j], v[j], red);
            red = fmaf(v[j + 1], v[j + 1], red);
            red = fmaf(v[j + 2], v[j + 2], red);
            red = fmaf(v[j + 3], v[j + 3], red);
          }
        } else {
          // u = m * dh; p and m rebuilt from the stored activation y = lrelu(p)
#pragma unroll
          for (int i = 0; i < HC / 8; ++i) {
            const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yraw[i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 yy = __bfloat1622float2(h2[e]);
              const int j = i * 8 + 2 * e;
              const float u0 = __uint_as_float(vr[j]) * scale * (yy.x > 0.f ? 1.f : slope);
              const float u1 = __uint_as_float(vr[j + 1]) * scale * (yy.y > 0.f ? 1.f : slope);
              red = fmaf(yy.x > 0.f ? yy.x : yy.x * inv_slope, u0, red);
              red = fmaf(yy.y > 0.f ? yy.y : yy.y * inv_slope, u1, red);
              v[j] = u0;
              v[j + 1] = u1;
            }
          }
        }
        if (NH == 2 && !ABW) {
          // columns HC..COUT-1 only contribute to the sum of squares in this pass
          uint32_t tb[2][16];
          tmem_ld<16>(t_addr + (uint32_t)HC, tb[0]);
#pragma unroll
          for (int c = 0; c < HC / 16; ++c) {
            tmem_ld_wait();
            if (c + 1 < HC / 16) tmem_ld<16>(t_addr + (uint32_t)(HC + (c + 1) * 16), tb[(c + 1) & 1]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = fmaf(__uint_as_float(tb[c & 1][j]), scale, bias_ptr[HC + c * 16 + j]);
              red = fmaf(a, a, red);
            }
          }
        }
        if (p.dbg & 2) {                     // experiment: accumulator drain only
          if (NH == 2 && mt == MT - 1) release_stage();
          continue;
        }
        float r = 1.f;
        if (!ABW && pn) r = rsqrtf(red * invC + 1e-8f);
        const float kpu = red * invC;        // ABW with PixelNorm: <p,u>/C
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          if (h == 1) {
            // ---- pass 2 (COUT = 128): re-read columns HC.. and normalise them
#pragma unroll
            for (int c = 0; c < HC; c += 32) tmem_ld<32>(t_addr + (uint32_t)(HC + c), vr + c);
            if constexpr (ABW) {
              const uint4 *yp = reinterpret_cast<const uint4 *>(p.y_prev + pix * COUT + HC);
#pragma unroll
              for (int i = 0; i < HC / 8; ++i) yraw[i] = __ldg(yp + i);
            }
            if (gt == 0) tma_store_wait_read0();     // first half's store has read the staging tile
            group_bar();
            tmem_ld_wait();
            if (mt == MT - 1) release_stage();
            if constexpr (!ABW) {
#pragma unroll
              for (int j = 0; j < HC; ++j)
                v[j] = fmaf(__uint_as_float(vr[j]), scale, bias_ptr[HC + j]);
            } else {
#pragma unroll
              for (int i = 0; i < HC / 8; ++i) {
                const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yraw[i]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 yy = __bfloat1622float2(h2[e]);
                  const int j = i * 8 + 2 * e;
                  v[j] = __uint_as_float(vr[j]) * scale * (yy.x > 0.f ? 1.f : slope);
                  v[j + 1] = __uint_as_float(vr[j + 1]) * scale * (yy.y > 0.f ? 1.f : slope);
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < HC / 8; ++i) {                       // 8 channels = 16 bytes per store
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = i * 8 + 2 * e;
              float a0, a1;
              if constexpr (!ABW) {
                a0 = v[j] * r;
                a1 = v[j + 1] * r;
                if (act) {
                  a0 = a0 > 0.f ? a0 : a0 * slope;
                  a1 = a1 > 0.f ? a1 : a1 * slope;
                }
              } else if (pn) {   // da = r (u - p <p,u>/C)
                const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&yraw[i]);
                const float2 yy = __bfloat1622float2(h2[e]);
                a0 = rp * fmaf(-(yy.x > 0.f ? yy.x : yy.x * inv_slope), kpu, v[j]);
                a1 = rp * fmaf(-(yy.y > 0.f ? yy.y : yy.y * inv_slope), kpu, v[j + 1]);
                v[j] = a0;
                v[j + 1] = a1;
              } else {
                a0 = v[j];
                a1 = v[j + 1];
              }
              __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
              pk[e] = *reinterpret_cast<uint32_t *>(&hh);
            }
            const uint32_t off = (uint32_t)row * (uint32_t)row_b + (uint32_t)i * 16u;
            *reinterpret_cast<uint4 *>(stage_ptr + swz(off, swz_bits)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          if (h == 0 && !ABW && pn) p.r_out[pix] = r;
          fence_proxy_async_smem();
          group_bar();
          if (gt == 0 && !(p.dbg & 1)) {
            tma_store_4d(&tmap_y, stage_s, h * HC, w0, h0, n);
            tma_store_commit();
          }
          if (!ABW && p.pool) {
            // 2x2 average pool of the staged tile (the reference's bilinear x0.5,
            // progan_modules.py:299): 32 pooled pixels x HC/8 16-byte chunks over 128 threads
            constexpr int CH8 = HC / 8;
            for (int item = gt; item < 32 * CH8; item += 128) {
              const int pp = item / CH8, c8 = item - pp * CH8;
              const int ph_ = pp >> 2, pw = pp & 3;
              const uint32_t cb = (uint32_t)c8 * 16u;
              float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const int srow = (2 * ph_ + (q4 >> 1)) * 8 + 2 * pw + (q4 & 1);
                const uint4 raw = *reinterpret_cast<const uint4 *>(
                    stage_ptr + swz((uint32_t)srow * (uint32_t)row_b + cb, swz_bits));
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
                __nv_bfloat162 hh = __floats2bfloat162_rn(0.25f * acc8[2 * e], 0.25f * acc8[2 * e + 1]);
                pk[e] = *reinterpret_cast<uint32_t *>(&hh);
              }
              *reinterpret_cast<uint4 *>(pool_g + swz((uint32_t)pp * (uint32_t)row_b + cb, swz_bits)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async_smem();
            group_bar();
            if (gt == 0 && !(p.dbg & 1)) {
              tma_store_4d(&tmap_yp, pool_s, h * HC, w0 >> 1, h0 >> 1, n);
              tma_store_commit();
            }
          }
          if (ABW && p.colsum != nullptr) {
            // bias gradient: per-channel sum of this half's da over the warp's 32 pixels by a
            // halving butterfly - each step trades half of the remaining channels with the partner
            // lane (HC - 1 shuffles in total)
            int nrem = HC;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              if (nrem >= 2) {
                const int hn = nrem / 2;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < HC / 2; ++j) {
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
            csum[2 * h] += v[0];
            if (HC >= 64) csum[2 * h + 1] += v[1];
          }
        }
      }
    }
    if (ABW && p.colsum != nullptr) {
      // channel owned by this lane after the butterfly (bit b of the lane picked the upper half
      // at the step with offset 2^b)
      int ch = 0;
      int hn = HC / 2;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        if (hn >= 1) {
          if (lane & off) ch += hn;
          hn >>= 1;
        }
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        atomicAdd(p.colsum + h * HC + ch, csum[2 * h]);
        if (HC >= 64) atomicAdd(p.colsum + h * HC + ch + 1, csum[2 * h + 1]);
      }
    }
    if (gt == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA exits while a peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_for(p.num_super) ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = CL;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = CL > 1 ? 2 : 1;
  if (max_ctas == 0) {
    max_ctas = sm_count();
    if (CL > 1) {
      int ncl = 0;
      cfg.gridDim = dim3((unsigned)(sm_count() / CL * CL));
      cfg.attrs = at + 1;          // the occupancy query takes the cluster attribute only
      cfg.numAttrs = 1;
      const bool occ_ok = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) == cudaSuccess && ncl > 0;
      cfg.attrs = at;
      cfg.numAttrs = 2;
      if (occ_ok)
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
  // experiment knobs, read ONCE per process (PG_DBG: epilogue / load / MMA switches for
  // profiles/r2/diag_conv4.py; PG_C4_RES / PG_C4_MT / PG_C4_CL: force a kernel variant)
  struct Knobs {
    int dbg = 0, res = -1, mt = -1, cl = -1, pair = 1;
    bool live = false;          // PG_DBG_LIVE=1: re-read PG_DBG on every launch (diag script)
    Knobs() {
      if (const char *e = getenv("PG_DBG")) dbg = atoi(e);
      if (const char *e = getenv("PG_C4_RES")) res = atoi(e);
      if (const char *e = getenv("PG_C4_MT")) mt = atoi(e);
      if (const char *e = getenv("PG_C4_CL")) cl = atoi(e);
      if (const char *e = getenv("PG_C4_PAIR")) pair = atoi(e);     // 0: no CTA-pair (cta_group::2) variants
      if (const char *e = getenv("PG_DBG_LIVE")) live = atoi(e) != 0;
    }
  };
  static const Knobs knobs;
  if (H % 16 || H < 16 || W % 8 || !(Cout == 32 || Cout == 64 || Cout == 128) || Cin % 32)
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
  p.dbg = knobs.dbg;
  if (knobs.live)
    if (const char *e = getenv("PG_DBG")) p.dbg = atoi(e);
  const int hc = Cout > 64 ? 64 : Cout;      // channels per staging tile; one tile per epilogue group
  const int nsb = Cout > 64 ? 1 : 2;         // staging tiles per group
  const int out_bytes = 2 * nsb * 128 * hc * 2 + (y_pool ? 2 * nsb * 32 * hc * 2 : 0);   // staging (+ pooled tiles)
  const int misc = 1024 + 8 * (2 * tc::kC4MaxA + 2 * tc::kC4MaxW + 5) + 16 + 16 + 128 * 4 + 2 * 128 * 4 + 64;
  const int budget = 227 * 1024 - out_bytes - misc;
  const int wres = 9 * p.ncb * p.wtile_bytes;
  const int force_res = knobs.res, force_mt = knobs.mt, force_cl = knobs.cl;
  int res = (wres + 2 * p.box_pad <= budget) ? 1 : 0;
  if (force_res >= 0 && (force_res == 0 || wres + 2 * p.box_pad <= budget)) res = force_res;
  if (BK == 32 && !res) return PG_ERR_UNSUPPORTED;
  if (!res && Cout < 64) return PG_ERR_UNSUPPORTED;
  int MT, CL;
  // CTA pair (M = 256 MMAs, half of the weights resident in each CTA) wherever it fits
  // ... and where the MMA phase of a tile is long enough to hide the cross-CTA handshakes of a
  // pair (tfull multicast -> peer epilogue -> remote tempty arrive: ~2x the latency of the local
  // protocol with the same two accumulator stages).  Measured (profiles/r2/diag_pair*.txt): K = 9 x 128
  // layers gain 7-9 %, the Cin <= 64 layers at 128 px LOSE 3-30 %; PG_C4_PAIR=2 forces pairs everywhere.
  const bool pair = knobs.pair != 0 && force_res != 0 && p.num_tiles % 2 == 0 &&
                    wres / 2 + 2 * p.box_pad <= budget && (Cin >= 128 || knobs.pair == 2);
  if (pair) {
    res = 1;
    CL = 2;
    MT = (p.num_tiles % 4 == 0 && 4 * Cout <= 512 && wres / 2 + 4 * p.box_pad <= budget) ? 2 : 1;
    if (force_mt == 1) MT = 1;
    p.a_stages = (budget - wres / 2) / (MT * p.box_pad);
    if (p.a_stages > tc::kC4MaxA) p.a_stages = tc::kC4MaxA;
    p.w_stages = 0;
  } else if (res) {
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
  const size_t smem = (size_t)(pair ? wres / 2 : (res ? wres : p.w_stages * p.wtile_bytes)) +
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
  PG_C4_TRY(BK_, NCB_, 1, true, CO_, 1) PG_C4_TRY(BK_, NCB_, 2, true, CO_, 1)                    \
  PG_C4_TRY(BK_, NCB_, 1, true, CO_, 2) PG_C4_TRY(BK_, NCB_, 2, true, CO_, 2)
#define PG_C4_STREAM(NCB_, CO_)                                                                  \
  PG_C4_TRY(64, NCB_, 1, false, CO_, 1) PG_C4_TRY(64, NCB_, 1, false, CO_, 2)                    \
  PG_C4_TRY(64, NCB_, 1, false, CO_, 4) PG_C4_TRY(64, NCB_, 2, false, CO_, 1)                    \
  PG_C4_TRY(64, NCB_, 2, false, CO_, 2) PG_C4_TRY(64, NCB_, 2, false, CO_, 4)
  PG_C4_RESIDENT(64, 1, 32) PG_C4_RESIDENT(64, 1, 64) PG_C4_RESIDENT(64, 1, 128)
  PG_C4_RESIDENT(64, 2, 32) PG_C4_RESIDENT(64, 2, 64)
  PG_C4_TRY(64, 2, 1, true, 128, 2) PG_C4_TRY(64, 2, 2, true, 128, 2)     /* 128 -> 128: pair only */
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
