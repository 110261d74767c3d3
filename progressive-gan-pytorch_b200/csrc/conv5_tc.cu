// 3x3 pad-1 implicit-GEMM convolution on CTA PAIRS (tcgen05.mma.cta_group::2).
//
// conv4_tc.cu's 128-channel layers are bound by shared-memory operand bandwidth: an M128 x N128
// x K16 MMA reads 4 KB of A and 4 KB of B in its 64 cycles = 128 B/clk, everything the SM has,
// so TMA fills and epilogue staging stall the pipe (measured ~65 % tensor-active).  A CTA pair
// issues ONE M = 256 MMA over two SMs: each CTA feeds its own 128-pixel tile (A) and only HALF
// of the weight rows (B), i.e. 96 B/clk, and half of the packed weights per CTA is small enough
// to stay RESIDENT in smem for every layer of the network (Cin = Cout = 128: 144 KB), which
// removes the weight stream altogether.
//
// Same tiling as conv4 (one 10x18 halo box per channel block, taps = start-address offsets).
// Protocol (rank 0 = leader issues all MMAs; both CTAs run producer + epilogue for their tile):
//   afull[s]  (leader)  count 2: each CTA's producer arms it with its box bytes, both TMA loads
//                       (.cta_group::2) complete_tx on the leader's barrier
//   aempty[s] (each)    count 1: leader's tcgen05.commit.cta_group::2 multicast
//   tfull[a]  (each)    count 1: same
//   tempty[a] (leader)  count 16: one arrive per epilogue warp of both CTAs (peer: remote arrive)
#include "tc_common.cuh"
#include <stdlib.h>

namespace pg {
namespace tc {

struct Conv5Params {
  int N, H, W, Cin;
  int tiles_w, tiles_h, num_tiles;
  int a_stages;
  int tmem_cols;
  int epi;
  float scale, slope;
  const float *bias;
  float *r_out;
  __nv_bfloat16 *y;
  int dbg;
};

constexpr int kC5Threads = 384;
constexpr int kC5MaxA = 8;

template <int BK, int NCB, int COUT>
__global__ void __launch_bounds__(kC5Threads, 1)
conv5_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_y, const Conv5Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t row_bytes = BK * 2u;
  constexpr uint32_t box_real = 18u * 10u * row_bytes;
  constexpr uint32_t kBoxPad = (box_real + 1023u) / 1024u * 1024u;
  constexpr uint32_t whalf_bytes = (uint32_t)(COUT / 2) * row_bytes;       // one tap tile, this CTA's rows
  constexpr uint32_t w_bytes = 9u * NCB * whalf_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_w = base;
  const uint32_t smem_a = base + w_bytes;
  const uint32_t bar_base = smem_a + (uint32_t)p.a_stages * kBoxPad;
  auto afull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bar_base + 8u * (uint32_t)(kC5MaxA + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kC5MaxA + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kC5MaxA + 2 + a); };
  const uint32_t wres_bar = bar_base + 8u * (uint32_t)(2 * kC5MaxA + 4);
  const uint32_t tmem_slot = wres_bar + 8u;
  const uint32_t bias_s = (tmem_slot + 16u + 15u) & ~15u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gbase + (tmem_slot - base));
  float *bias_ptr = reinterpret_cast<float *>(gbase + (bias_s - base));
  float *ss_buf = bias_ptr + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pid = (int)cluster_id_x();
  const int npairs = (int)gridDim.x / 2;
  const int num_sp = p.num_tiles / 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(afull(s), 2);
      mbar_init(aempty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 16);
    }
    mbar_init(wres_bar, 2);
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < COUT; c += kC5Threads) bias_ptr[c] = p.bias ? p.bias[c] : 0.f;
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers exist before anything is signalled
  if (warp == 2) tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== activation producer (both CTAs) =====================
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      for (int sp = pid; sp < num_sp; sp += npairs) {
        const int tile = 2 * sp + (int)rank;
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          mbar_wait(aempty(as), aph ^ 1u);
          if (leader) mbar_expect_tx(afull(as), box_real);
          else mbar_expect_tx_remote(mapa_shared(afull(as), 0), box_real);
          tma_load_4d_pair(smem_a + (uint32_t)as * kBoxPad, &tmap_x, afull(as), cb * BK, tw * 8 - 1,
                           th * 16 - 1, n);
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weights: this CTA's half of every tap tile, once =====================
    if (lane == 0) {
      if (leader) mbar_expect_tx(wres_bar, w_bytes);
      else mbar_expect_tx_remote(mapa_shared(wres_bar, 0), w_bytes);
      for (int tap = 0; tap < 9; ++tap)
        for (int cb = 0; cb < NCB; ++cb)
          tma_load_2d_pair(smem_w + (uint32_t)(tap * NCB + cb) * whalf_bytes, &tmap_w, wres_bar,
                           tap * p.Cin + cb * BK, (int)rank * (COUT / 2));
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader) {
      constexpr uint32_t layout = row_bytes == 128 ? 2u : 4u;
      constexpr uint32_t sbo_a = 10u * row_bytes;
      constexpr uint32_t sbo_b = 8u * row_bytes;
      constexpr int nk = BK / 16;
      constexpr uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      constexpr uint32_t hi_b = ((sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      constexpr uint32_t lbo_lo = 1u << 16;
      constexpr uint32_t whalf16 = whalf_bytes >> 4;
      const uint32_t idesc = make_idesc_bf16(256, COUT, 0, 0);
      mbar_wait(wres_bar, 0);
      tc_fence_after();
      int as = 0, acc = 0;
      uint32_t aph = 0, acc_phase = 0;
      const bool prof = (p.dbg & 64) != 0;
      long long w_t = 0, w_a = 0, t_start = clock64();
      int ntile = 0;
      for (int sp = pid; sp < num_sp; sp += npairs) {
        long long c0 = prof ? clock64() : 0;
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (prof) w_t += clock64() - c0;
        ++ntile;
        const uint32_t d_base = tmem_base + (uint32_t)(acc * COUT);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          c0 = prof ? clock64() : 0;
          mbar_wait(afull(as), aph);
          tc_fence_after();
          if (prof) w_a += clock64() - c0;
          const uint32_t a16 = ((smem_a + (uint32_t)as * kBoxPad) >> 4) | lbo_lo;
          const uint32_t w16 = (smem_w >> 4) | lbo_lo;
          if (elect_one_sync()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dh = tap / 3, dw = tap % 3;
              const uint32_t b16 = w16 + (uint32_t)(tap * NCB + cb) * whalf16;
              const uint32_t a_off16 = ((uint32_t)(dh * 10 + dw) * row_bytes) >> 4;
#pragma unroll
              for (int k = 0; k < nk; ++k) {
                const uint64_t ad = ((uint64_t)hi_a << 32) | (uint64_t)(a16 + a_off16 + (uint32_t)(k * 2));
                const uint64_t bd = ((uint64_t)hi_b << 32) | (uint64_t)(b16 + (uint32_t)(k * 2));
                umma_bf16_2cta(d_base, ad, bd, idesc, (cb | tap | k) ? 1u : 0u);
              }
            }
            umma_commit2(aempty(as));
            if (cb == NCB - 1) umma_commit2(tfull(acc));
          }
          __syncwarp();
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (prof && blockIdx.x == 0 && lane == 0)
        printf("conv5 mma warp: %d tiles, total %lld cycles, wait tempty %lld, wait afull %lld\n", ntile,
               clock64() - t_start, w_t, w_a);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, 8 warps each) =====================
    constexpr int CPT = COUT / 2;
    constexpr int out_chunk = (COUT % 64 == 0) ? 64 : 32;
    constexpr int chunk_rows_bytes = out_chunk * 2;
    constexpr int swz_bits = chunk_rows_bytes == 128 ? 3 : 2;
    constexpr int n_chunks = COUT / out_chunk;
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int col0 = part * CPT;
    const float invC = 1.f / (float)COUT;
    const float scale = p.scale, slope = p.slope;
    const uint32_t tempty_leader0 = leader ? tempty(0) : mapa_shared(tempty(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sp = pid; sp < num_sp; sp += npairs) {
      const int tile = 2 * sp + (int)rank;
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int n = tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * 8, h0 = th * 16;
      const long long pix = ((long long)n * p.H + (h0 + (row >> 3))) * p.W + w0 + (row & 7);
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * COUT + col0);
      uint32_t vr[CPT];
      tmem_ld<CPT>(t_addr, vr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                       // accumulator stage drained by this warp
        if (leader) mbar_arrive(tempty_leader0 + 8u * (uint32_t)acc);
        else mbar_arrive_remote(tempty_leader0 + 8u * (uint32_t)acc);
      }
      float v[CPT];
      float r = 1.f;
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
      if (p.epi == PG_EPI_PN_LRELU) {
        ss_buf[part * 128 + row] = ss;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        r = rsqrtf((ss_buf[row] + ss_buf[128 + row]) * invC + 1e-8f);
      }
      // direct stores: this thread owns CPT consecutive channels of one pixel (CPT*2 contiguous
      // bytes).  No smem staging tile: its 32 KB buy the third activation stage that hides the
      // ~4400-cycle refill round trip of a CTA pair (measured), and the MMA phase of a
      // 128-channel tile (4608 cycles) leaves the LSU time for the scattered 16-byte stores.
      {
        uint4 *dst = reinterpret_cast<uint4 *>(p.y + pix * COUT + col0);
#pragma unroll
        for (int i = 0; i < CPT / 8; ++i) {
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
          dst[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if (p.epi == PG_EPI_PN_LRELU && part == 0) p.r_out[pix] = r;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no CTA exits (or frees TMEM) while its peer may still use it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int BK, int NCB, int COUT>
static int launch_c5(const void *x, const void *wp, void *y, Conv5Params p, int N, int H, int W, int Cin,
                     cudaStream_t stream) {
  constexpr int box_pad = (18 * 10 * BK * 2 + 1023) / 1024 * 1024;
  constexpr int w_bytes = 9 * NCB * (COUT / 2) * BK * 2;
  constexpr int out_bytes = 0;             // the epilogue stores straight from registers
  const int misc = 1024 + 8 * (2 * kC5MaxA + 5) + 16 + 16 + 128 * 4 + 2 * 128 * 4 + 64;
  int a_stages = (227 * 1024 - w_bytes - out_bytes - misc) / box_pad;
  if (a_stages > kC5MaxA) a_stages = kC5MaxA;
  if (a_stages < 2) return PG_ERR_UNSUPPORTED;
  p.a_stages = a_stages;
  int cols = 2 * COUT;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const size_t smem = (size_t)w_bytes + (size_t)a_stages * box_pad + out_bytes + misc;
  CUtensorMap tx, tw_, ty;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)BK, 10u, 18u, 1u};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, BK * 2, "pg_conv_tc/v5(x)")) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)COUT};
    uint64_t str[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(COUT / 2)};
    if (int rc = make_tmap_bf16(&tw_, wp, 2, dims, str, box, BK * 2, "pg_conv_tc/v5(w)")) return rc;
  }
  constexpr int out_chunk = (COUT % 64 == 0) ? 64 : 32;
  {
    uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)COUT * 2, (uint64_t)W * COUT * 2, (uint64_t)H * W * COUT * 2};
    uint32_t box[4] = {(uint32_t)out_chunk, 8u, 16u, 1u};
    if (int rc = make_tmap_bf16(&ty, y, 4, dims, str, box, out_chunk * 2, "pg_conv_tc/v5(y)")) return rc;
  }
  auto kern = conv5_tc_kernel<BK, NCB, COUT>;
  static bool attr_set = false;
  static int max_ctas = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("pg_conv_tc/v5: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kC5Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_ctas == 0) {
    int ncl = 0;
    cfg.gridDim = dim3((unsigned)(sm_count() / 2 * 2));
    if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) == cudaSuccess && ncl > 0) max_ctas = ncl * 2;
    else { (void)cudaGetLastError(); max_ctas = sm_count() / 2 * 2; }
    if (max_ctas > sm_count()) max_ctas = sm_count() / 2 * 2;
  }
  int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  grid = grid / 2 * 2;
  cfg.gridDim = dim3((unsigned)grid);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw_, ty, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pg_conv_tc/v5: launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace tc

// PG_ERR_UNSUPPORTED (no error set) when the shape is not served by the pair kernel.
int conv5_tc_launch(const void *x, const void *wp, const float *bias, void *y, float *r_out, int N,
                    int H, int W, int Cin, int Cout, float scale, int epi, float slope,
                    cudaStream_t stream, void *y_pool) {
  int on = 0;
  if (const char *e = getenv("PG_CONV_V5")) on = atoi(e);
  if (!on || y_pool) return PG_ERR_UNSUPPORTED;
  if (H % 16 || H < 16 || W % 8) return PG_ERR_UNSUPPORTED;
  tc::Conv5Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * N;
  if (p.num_tiles % 2) return PG_ERR_UNSUPPORTED;
  p.epi = epi; p.scale = scale; p.slope = slope; p.bias = bias; p.r_out = r_out;
  p.y = (__nv_bfloat16 *)y;
  p.dbg = 0;
  if (const char *e = getenv("PG_DBG")) p.dbg = atoi(e);
  int bk = 64;
  if (const char *e = getenv("PG_C5_BK")) bk = atoi(e);
  if (Cin == 128 && Cout == 128 && bk == 32) return tc::launch_c5<32, 4, 128>(x, wp, y, p, N, H, W, Cin, stream);
  if (Cin == 128 && Cout == 128) return tc::launch_c5<64, 2, 128>(x, wp, y, p, N, H, W, Cin, stream);
  if (Cin == 64 && Cout == 128) return tc::launch_c5<64, 1, 128>(x, wp, y, p, N, H, W, Cin, stream);
  if (on >= 2) {
    if (Cin == 128 && Cout == 64) return tc::launch_c5<64, 2, 64>(x, wp, y, p, N, H, W, Cin, stream);
    if (Cin == 64 && Cout == 64) return tc::launch_c5<64, 1, 64>(x, wp, y, p, N, H, W, Cin, stream);
  }
  return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
