// 3x3 pad-1 weight-gradient, second generation ("column-halo") kernel.
//
// Same math as wgrad_tc.cu (dW[co][ci][tap] = sum_pix dy[pix,co] x[pix+tap,ci], K = pixels,
// MN-major operands straight from NHWC, accumulators resident in TMEM for the CTA's whole pixel
// range) but — like conv3_tc.cu — the activation is fetched as THREE 8x18-pixel boxes per
// 64-channel block (one per horizontal tap offset dw) instead of nine shifted 128-pixel tiles:
// the vertical taps dh are 1024-byte-aligned start offsets into the same smem box, i.e. K-window
// shifts of whole tile rows, so the MN-major UMMA descriptors stay canonical.  L2 lines per
// pixel tile drop from 9*128 + dy to 3*144 + dy.
//
// M stacking inside one dw box: M = 128 = n_atoms_m atoms of 64 (or 32) channels
//   Cin =  32: atoms = dh 0,1,2 (+1 dummy), LBO = one tile row
//   Cin =  64: atoms = {dh0,dh1}, {dh2,dummy}: two groups per dw, LBO = one tile row
//   Cin = 128: atoms = the two 64-channel blocks of one tap, LBO = box size, three groups per dw
// Accumulator groups that do not fit 512 TMEM columns are split over blockIdx.y by dw.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pg {
namespace tc {

struct Wgrad3Params {
  int N, H, W, Cin, Cout;
  int tiles_w, tiles_h, num_tiles;
  int atomM, atomN, n_atoms_m, n_atoms_n, atoms_per_tap, ncb;
  int tpg;                 // taps (dh values) per accumulator group
  int gpd;                 // groups per dw
  int dw_per_pass;         // dw values handled by one blockIdx.y
  int units;               // x ring depth (unit = all channel blocks of one dw box)
  int box_bytes;           // one 8x18 box of atomM channels
  int unit_bytes;          // ncb * box_bytes
  int dy_atom_bytes, dy_bytes;
  int lbo_bytes;           // atom stride inside a group
  int group_stride, tmem_cols;
  float *dwp;              // [9][Cout][Cin] fp32 workspace (zeroed)
};

constexpr int kW3Threads = 256;

__global__ void __launch_bounds__(kW3Threads, 1)
wgrad3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                 const Wgrad3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x0 = base;
  // one spare box after the ring: dummy atoms of the last unit read (and ignore) it
  const uint32_t smem_dy0 = base + (uint32_t)(p.units * p.unit_bytes + p.box_bytes);
  const uint32_t bar_base = smem_dy0 + 2u * (uint32_t)p.dy_bytes;
  auto xfull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto xempty = [&](int s) { return bar_base + 8u * (uint32_t)(p.units + s); };
  auto dyfull = [&](int d) { return bar_base + 8u * (uint32_t)(2 * p.units + d); };
  auto dyempty = [&](int d) { return bar_base + 8u * (uint32_t)(2 * p.units + 2 + d); };
  const uint32_t done_bar = bar_base + 8u * (uint32_t)(2 * p.units + 4);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dw0 = blockIdx.y * p.dw_per_pass;
  const int dw1 = min(dw0 + p.dw_per_pass, 3);
  const int ndw = dw1 - dw0;

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
        mbar_expect_tx(dyfull(ds), (uint32_t)p.dy_bytes);
        for (int a = 0; a < p.n_atoms_n; ++a)
          tma_load_4d(smem_dy0 + (uint32_t)(ds * p.dy_bytes + a * p.dy_atom_bytes), &tmap_dy,
                      dyfull(ds), a * p.atomN, w0, h0, n);
        if (++ds == 2) {
          ds = 0;
          dphase ^= 1u;
        }
        for (int dwi = dw0; dwi < dw1; ++dwi) {
          mbar_wait(xempty(xs), xphase ^ 1u);
          mbar_expect_tx(xfull(xs), (uint32_t)p.unit_bytes);
          for (int cb = 0; cb < p.ncb; ++cb)
            tma_load_4d(smem_x0 + (uint32_t)(xs * p.unit_bytes + cb * p.box_bytes), &tmap_x, xfull(xs),
                        cb * p.atomM, w0 + dwi - 1, h0 - 1, n);
          if (++xs == p.units) {
            xs = 0;
            xphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, p.Cout, 1, 1);   // both operands MN-major
    const uint32_t rowA = (uint32_t)p.atomM * 2u, rowB = (uint32_t)p.atomN * 2u;
    const uint32_t layA = rowA == 128 ? 2u : 4u, layB = rowB == 128 ? 2u : 4u;
    const uint32_t hiA = (((8u * rowA) >> 4) & 0x3FFFu) | (1u << 14) | (layA << 29);
    const uint32_t hiB = (((8u * rowB) >> 4) & 0x3FFFu) | (1u << 14) | (layB << 29);
    const uint32_t lboA = (((uint32_t)p.lbo_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t lboB = (((uint32_t)p.dy_atom_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t stepA = (16u * rowA) >> 4, stepB = (16u * rowB) >> 4;   // 16 pixels = 2 tile rows
    // group g of a dw box starts at tile row g*tpg (when the atoms are successive dh windows)
    const uint32_t gstep = (p.atoms_per_tap == 1) ? ((uint32_t)p.tpg * 8u * rowA) >> 4 : (8u * rowA) >> 4;
    int xs = 0, ds = 0;
    uint32_t xphase = 0, dphase = 0;
    uint32_t first = 1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(dyfull(ds), dphase);
      tc_fence_after();
      const uint32_t b_lo = ((smem_dy0 + (uint32_t)(ds * p.dy_bytes)) >> 4) | lboB;
      for (int d = 0; d < ndw; ++d) {
        mbar_wait(xfull(xs), xphase);
        tc_fence_after();
        const uint32_t a_unit = ((smem_x0 + (uint32_t)(xs * p.unit_bytes)) >> 4) | lboA;
        if (elect_one_sync()) {
          for (int g = 0; g < p.gpd; ++g) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((d * p.gpd + g) * p.group_stride);
            const uint32_t a_lo = a_unit + (uint32_t)g * gstep;
#pragma unroll
            for (int k = 0; k < 8; ++k) {          // 128 pixels = 8 x K16
              const uint64_t ad = ((uint64_t)hiA << 32) | (uint64_t)(a_lo + (uint32_t)k * stepA);
              const uint64_t bd = ((uint64_t)hiB << 32) | (uint64_t)(b_lo + (uint32_t)k * stepB);
              umma_bf16(d_tmem, ad, bd, idesc, k ? 1u : (first ^ 1u));
            }
          }
          umma_commit(xempty(xs));
          if (d == ndw - 1) umma_commit(dyempty(ds));
        }
        __syncwarp();
        if (++xs == p.units) {
          xs = 0;
          xphase ^= 1u;
        }
      }
      if (++ds == 2) {
        ds = 0;
        dphase ^= 1u;
      }
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
      const int a = m / p.atomM;                     // atom of this lane
      for (int d = 0; d < ndw; ++d) {
        for (int g = 0; g < p.gpd; ++g) {
          const int dh = g * p.tpg + a / p.atoms_per_tap;
          const int tap = dh * 3 + (dw0 + d);
          const int ci = (a % p.atoms_per_tap) * p.atomM + m % p.atomM;
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                  (uint32_t)((d * p.gpd + g) * p.group_stride);
          for (int c0 = 0; c0 < p.Cout; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_addr + (uint32_t)c0, v);   // warp-collective
            tmem_ld_wait();
            if (dh < 3) {
              float *dst = p.dwp + ((size_t)tap * p.Cout + c0) * p.Cin + ci;
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * p.Cin, __uint_as_float(v[j]));
            }
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

}  // namespace tc

// PG_ERR_UNSUPPORTED (no error set) when the shape is not eligible -> first-generation kernel.
// The workspace must already be zeroed; the caller runs the unpack kernel afterwards.
int wgrad3_tc_launch(const void *x, const void *dy, float *workspace, int N, int H, int W, int Cin,
                     int Cout, cudaStream_t stream) {
  if (const char *e = getenv("PG_WGRAD_V2"))
    if (atoi(e) == 0) return PG_ERR_UNSUPPORTED;
  if (H % 16 || H < 32 || W % 8) return PG_ERR_UNSUPPORTED;
  if (!(Cin == 32 || Cin == 64 || Cin == 128) || !(Cout == 32 || Cout == 64 || Cout == 128))
    return PG_ERR_UNSUPPORTED;
  tc::Wgrad3Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.num_tiles = p.tiles_w * p.tiles_h * N;
  p.atomM = Cin >= 64 ? 64 : 32;
  p.atomN = Cout >= 64 ? 64 : 32;
  p.n_atoms_m = 128 / p.atomM;
  p.n_atoms_n = Cout / p.atomN;
  p.atoms_per_tap = Cin / p.atomM;              // 1, 1, 2
  p.ncb = p.atoms_per_tap;
  p.tpg = p.n_atoms_m / p.atoms_per_tap;        // 4, 2, 1
  p.gpd = (3 + p.tpg - 1) / p.tpg;              // 1, 2, 3
  p.box_bytes = 18 * 8 * p.atomM * 2;
  p.unit_bytes = p.ncb * p.box_bytes;
  p.dy_atom_bytes = 128 * p.atomN * 2;
  p.dy_bytes = p.n_atoms_n * p.dy_atom_bytes;
  p.lbo_bytes = (p.atoms_per_tap == 1) ? 8 * p.atomM * 2 : p.box_bytes;
  p.group_stride = Cout;
  const int max_groups = 512 / Cout;
  p.dw_per_pass = max_groups / p.gpd;
  if (p.dw_per_pass > 3) p.dw_per_pass = 3;
  if (p.dw_per_pass < 1) return PG_ERR_UNSUPPORTED;
  const int passes = (3 + p.dw_per_pass - 1) / p.dw_per_pass;
  int cols = p.dw_per_pass * p.gpd * Cout;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.dwp = workspace;
  const int misc = 1024 + 8 * (2 * 10 + 5) + 64;
  int units = (227 * 1024 - 2 * p.dy_bytes - p.box_bytes - misc) / p.unit_bytes;
  if (units > 10) units = 10;
  if (units < 2) return PG_ERR_UNSUPPORTED;
  p.units = units;
  const size_t smem = (size_t)units * p.unit_bytes + p.box_bytes + 2 * p.dy_bytes + misc;

  CUtensorMap tx, tdy;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.atomM, 8u, 18u, 1u};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, p.atomM * 2, "pg_conv_wgrad_tc/v2(x)")) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {(uint32_t)p.atomN, 8u, 16u, 1u};
    if (int rc = make_tmap_bf16(&tdy, dy, 4, dims, str, box, p.atomN * 2, "pg_conv_wgrad_tc/v2(dy)")) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::wgrad3_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("pg_conv_wgrad_tc/v2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    attr_set = true;
  }
  int gx = sm_count() / passes;
  if (gx > p.num_tiles) gx = p.num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, passes);
  tc::wgrad3_tc_kernel<<<grid, tc::kW3Threads, smem, stream>>>(tx, tdy, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pg_conv_wgrad_tc/v2: CUDA launch failed: %s", cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace pg
