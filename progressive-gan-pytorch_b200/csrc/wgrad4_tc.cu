// 3x3 pad-1 weight-gradient, third generation ("single halo box") kernel.
//
// dW[tap][co][ci] = sum_pix dy[pix,co] x[pix+tap,ci], K = pixels,
// both operands MN-major straight from NHWC, accumulators resident in TMEM for the CTA's whole
// pixel range.  Like conv4_tc.cu the activation is fetched as ONE 10(w) x 18(h) pixel box per
// channel block and tile; every tap (dh,dw) is the start-address offset (dh*10 + dw) rows into
// it (the swizzle is a function of the absolute smem address, verified on B200), a K step of 16
// pixels is two tile rows 10 box rows apart (SBO = 10 rows).  Since any two taps are just two
// start addresses, the M = 128 stacking is free to pair them:
//   Cin =  64: groups = tap pairs (0,1)(2,3)(4,5)(6,7)(8,-): 5 groups, 10% idle rows (was 25%)
//   Cin =  32: groups = (dh; dw 0,1,2,-): LBO must be one constant, so 3 taps + 1 idle atom
//   Cin = 128: groups = single taps, atoms = the two 64-channel boxes (LBO = box stride)
// Groups that do not fit 512 TMEM columns are split over blockIdx.y ("passes").
// L2 lines per tile: 180 box rows (+ dy) instead of 3 x 144.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pg {
namespace tc {

struct Wgrad4Params {
  int N, H, W;
  int tiles_w, tiles_h, num_tiles;
  int units;                 // x ring depth (unit = all channel boxes of one tile)
  int gpp;                   // groups per pass
  int tmem_cols;
  float *dwp;                // [9][Cout][Cin] fp32 workspace (accumulated into)
};

constexpr int kW4Threads = 256;
constexpr int kW4MaxUnits = 8;

template <int CIN, int COUT>
__global__ void __launch_bounds__(kW4Threads, 1)
wgrad4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                 const Wgrad4Params p) {
  constexpr int ATOM_M = CIN >= 64 ? 64 : 32;
  constexpr int ATOM_N = COUT >= 64 ? 64 : 32;
  constexpr int NCB = CIN / ATOM_M;                       // 1, 1, 2
  constexpr int N_ATOMS_N = COUT / ATOM_N;
  constexpr uint32_t rowA = ATOM_M * 2u, rowB = ATOM_N * 2u;
  constexpr uint32_t box_real = 18u * 10u * rowA;
  constexpr uint32_t kBoxPad = (box_real + 1023u) / 1024u * 1024u;
  constexpr uint32_t unit_bytes = NCB * kBoxPad;
  constexpr uint32_t dy_atom_bytes = 128u * rowB;
  constexpr uint32_t dy_bytes = N_ATOMS_N * dy_atom_bytes;
  constexpr int TAPS_PER_GROUP = CIN == 64 ? 2 : (CIN == 32 ? 3 : 1);
  constexpr int NGROUPS = (9 + TAPS_PER_GROUP - 1) / TAPS_PER_GROUP;   // 5, 3, 9

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x0 = base;
  // one spare box after the ring: the idle atom of the last group reads (and ignores) past its box
  const uint32_t smem_dy0 = base + (uint32_t)p.units * unit_bytes + kBoxPad;
  const uint32_t bar_base = smem_dy0 + 2u * dy_bytes;
  auto xfull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto xempty = [&](int s) { return bar_base + 8u * (uint32_t)(kW4MaxUnits + s); };
  auto dyfull = [&](int d) { return bar_base + 8u * (uint32_t)(2 * kW4MaxUnits + d); };
  auto dyempty = [&](int d) { return bar_base + 8u * (uint32_t)(2 * kW4MaxUnits + 2 + d); };
  const uint32_t done_bar = bar_base + 8u * (uint32_t)(2 * kW4MaxUnits + 4);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = blockIdx.y * p.gpp;
  const int g1 = min(g0 + p.gpp, NGROUPS);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.units; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(dyfull(d), 1);
      mbar_init(dyempty(d), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // everything above touches only this CTA's shared memory / TMEM: it overlaps the tail of the
  // previous kernel; from here on the predecessors' results are read
  pg::grid_dep_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int xs = 0, ds = 0;
      uint32_t xphase = 0, dphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * 8, h0 = th * 16;
        mbar_wait(dyempty(ds), dphase ^ 1u);
        mbar_expect_tx(dyfull(ds), dy_bytes);
#pragma unroll
        for (int a = 0; a < N_ATOMS_N; ++a)
          tma_load_4d(smem_dy0 + (uint32_t)ds * dy_bytes + (uint32_t)a * dy_atom_bytes, &tmap_dy,
                      dyfull(ds), a * ATOM_N, w0, h0, n);
        if (++ds == 2) { ds = 0; dphase ^= 1u; }
        mbar_wait(xempty(xs), xphase ^ 1u);
        mbar_expect_tx(xfull(xs), NCB * box_real);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb)
          tma_load_4d(smem_x0 + (uint32_t)xs * unit_bytes + (uint32_t)cb * kBoxPad, &tmap_x, xfull(xs),
                      cb * ATOM_M, w0 - 1, h0 - 1, n);
        if (++xs == p.units) { xs = 0; xphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, COUT, 1, 1);   // both operands MN-major
    constexpr uint32_t layA = rowA == 128 ? 2u : 4u, layB = rowB == 128 ? 2u : 4u;
    // K groups (8 pixels = one tile row) are 10 box rows apart in x, 8 rows apart in the dy tile
    constexpr uint32_t hiA = (((10u * rowA) >> 4) & 0x3FFFu) | (1u << 14) | (layA << 29);
    constexpr uint32_t hiB = (((8u * rowB) >> 4) & 0x3FFFu) | (1u << 14) | (layB << 29);
    constexpr uint32_t lboB = ((dy_atom_bytes >> 4) & 0x3FFFu) << 16;
    constexpr uint32_t stepA = (20u * rowA) >> 4, stepB = (16u * rowB) >> 4;   // 16 pixels = 2 tile rows
    int xs = 0, ds = 0;
    uint32_t xphase = 0, dphase = 0;
    uint32_t first = 1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(dyfull(ds), dphase);
      mbar_wait(xfull(xs), xphase);
      tc_fence_after();
      const uint32_t b_lo = ((smem_dy0 + (uint32_t)ds * dy_bytes) >> 4) | lboB;
      const uint32_t a_unit = (smem_x0 + (uint32_t)xs * unit_bytes) >> 4;
      if (elect_one_sync()) {
#pragma unroll
        for (int g = 0; g < NGROUPS; ++g) {
          if (g >= g0 && g < g1) {
            // first tap of the group and the distance to the next stacked atom
            const int t0 = g * TAPS_PER_GROUP;
            const uint32_t off_rows = (uint32_t)((t0 / 3) * 10 + t0 % 3);
            uint32_t lbo_bytes;
            if (CIN == 128) lbo_bytes = kBoxPad;                       // second channel box
            else if (CIN == 32) lbo_bytes = rowA;                      // next dw
            else {                                                     // CIN == 64: next tap
              const int t1 = t0 + 1 < 9 ? t0 + 1 : t0;                 // idle atom of the last group
              const uint32_t off1 = (uint32_t)((t1 / 3) * 10 + t1 % 3);
              lbo_bytes = (t1 == t0 ? 1u : off1 - off_rows) * rowA;
            }
            const uint32_t a_lo = (a_unit + ((off_rows * rowA) >> 4)) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
            const uint32_t d_tmem = tmem_base + (uint32_t)((g - g0) * COUT);
#pragma unroll
            for (int k = 0; k < 8; ++k) {          // 128 pixels = 8 x K16
              const uint64_t ad = ((uint64_t)hiA << 32) | (uint64_t)(a_lo + (uint32_t)k * stepA);
              const uint64_t bd = ((uint64_t)hiB << 32) | (uint64_t)(b_lo + (uint32_t)k * stepB);
              umma_bf16(d_tmem, ad, bd, idesc, k ? 1u : (first ^ 1u));
            }
          }
        }
        umma_commit(xempty(xs));
        umma_commit(dyempty(ds));
      }
      __syncwarp();
      if (++xs == p.units) { xs = 0; xphase ^= 1u; }
      if (++ds == 2) { ds = 0; dphase ^= 1u; }
      first = 0;
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== final reduction =====================
    if ((int)blockIdx.x < p.num_tiles) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const int a = m / ATOM_M;                      // atom of this lane
      for (int g = g0; g < g1; ++g) {
        int tap, ci;
        bool live;
        if (CIN == 128) {
          tap = g; ci = a * 64 + m % 64; live = true;
        } else {
          tap = g * TAPS_PER_GROUP + a; ci = m % ATOM_M; live = a < TAPS_PER_GROUP && tap < 9;
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g0) * COUT);
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + (uint32_t)c0, v);   // warp-collective
          tmem_ld_wait();
          if (live) {
            float *dst = p.dwp + ((size_t)tap * COUT + c0) * CIN + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * CIN, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int CIN, int COUT>
static int launch_w4(const void *x, const void *dy, float *workspace, int N, int H, int W,
                     cudaStream_t stream) {
  constexpr int ATOM_M = CIN >= 64 ? 64 : 32;
  constexpr int ATOM_N = COUT >= 64 ? 64 : 32;
  constexpr int NCB = CIN / ATOM_M;
  constexpr int TPG = CIN == 64 ? 2 : (CIN == 32 ? 3 : 1);
  constexpr int NGROUPS = (9 + TPG - 1) / TPG;
  constexpr int box_pad = (18 * 10 * ATOM_M * 2 + 1023) / 1024 * 1024;
  constexpr int unit_bytes = NCB * box_pad;
  constexpr int dy_bytes = 128 * COUT * 2;
  Wgrad4Params p;
  p.N = N; p.H = H; p.W = W;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * N;
  const int max_groups = 512 / COUT;
  const int passes = (NGROUPS + max_groups - 1) / max_groups;
  p.gpp = (NGROUPS + passes - 1) / passes;
  int cols = p.gpp * COUT;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.dwp = workspace;
  const int misc = 1024 + 8 * (2 * kW4MaxUnits + 5) + 64;
  // leave ~32 KB of the SM's shared memory free: the kernel runs on a side stream next to the
  // bandwidth-bound activation-backward kernels, whose bias-gradient reduction needs a few KB
  int units = (195 * 1024 - 2 * dy_bytes - box_pad - misc) / unit_bytes;
  if (units > kW4MaxUnits) units = kW4MaxUnits;
  if (units < 2) return PG_ERR_UNSUPPORTED;
  p.units = units;
  const size_t smem = (size_t)units * unit_bytes + box_pad + 2 * dy_bytes + misc;

  CUtensorMap tx, tdy;
  {
    uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)CIN * 2, (uint64_t)W * CIN * 2, (uint64_t)H * W * CIN * 2};
    uint32_t box[4] = {(uint32_t)ATOM_M, 10u, 18u, 1u};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, ATOM_M * 2, "pg_conv_wgrad_tc/v4(x)")) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)COUT * 2, (uint64_t)W * COUT * 2, (uint64_t)H * W * COUT * 2};
    uint32_t box[4] = {(uint32_t)ATOM_N, 8u, 16u, 1u};
    if (int rc = make_tmap_bf16(&tdy, dy, 4, dims, str, box, ATOM_N * 2, "pg_conv_wgrad_tc/v4(dy)")) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad4_tc_kernel<CIN, COUT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("pg_conv_wgrad_tc/v4: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    attr_set = true;
  }
  int gx = sm_count() / passes;
  if (gx > p.num_tiles) gx = p.num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, passes);
  pg::launcher(wgrad4_tc_kernel<CIN, COUT>, grid, kW4Threads, smem, stream, p.num_tiles)(tx, tdy, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pg_conv_wgrad_tc/v4: CUDA launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace tc

// PG_ERR_UNSUPPORTED (no error set) when the shape is not eligible -> older kernels.
// The workspace is accumulated into; the caller runs the unpack kernel afterwards.
int wgrad4_tc_launch(const void *x, const void *dy, float *workspace, int N, int H, int W, int Cin,
                     int Cout, cudaStream_t stream) {
  if (H % 16 || H < 16 || W % 8) return PG_ERR_UNSUPPORTED;
#define PG_W4(CI, CO) \
  if (Cin == CI && Cout == CO) return tc::launch_w4<CI, CO>(x, dy, workspace, N, H, W, stream);
  PG_W4(32, 32) PG_W4(32, 64) PG_W4(32, 128)
  PG_W4(64, 32) PG_W4(64, 64) PG_W4(64, 128)
  PG_W4(128, 32) PG_W4(128, 64) PG_W4(128, 128)
#undef PG_W4
  return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
