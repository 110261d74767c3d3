// 3x3 pad-1 convolution weight-gradient on tcgen05 tensor cores.
//
// Replaces aten::convolution_backward (weight part; cuDNN wgrad) for the EqualConv2d 3x3
// layers (reference progan_modules.py:63-73) — used four times per conv per iteration:
// D real, D fake, and twice inside the WGAN-GP double backward (train.py:130,139,151).
//
//   dWl[co][ci][tap] = scale * sum_pix dy[pix, co] * x[pix + tap_offset, ci]
//
// GEMM view: K = pixels (reduction), both operands are "MN-major" straight out of NHWC:
//   A (M side) = x tap tiles, stacked so that M = 128 = (128/Cin) taps x Cin channels
//   B (N side) = dy tile, N = Cout
//   D          = [128 (tap,ci) lanes] x [Cout columns] fp32 per tap group, in TMEM.
// The accumulators stay resident in TMEM for the CTA's whole pixel range (no epilogue per
// tile); tap groups that do not fit 512 TMEM columns are split over blockIdx.y ("passes").
// Tiles are TMA 4-D boxes (channel atom of 64 or 32, bw x bh x bn pixels) with 128B/64B
// swizzle; out-of-bounds pixels of the shifted x boxes are zero-filled == conv padding.
// At the end each CTA reduces its partial into a [tap][co][ci] fp32 workspace with
// coalesced red.global.add.f32 (a warp's 32 lanes = 32 consecutive ci), and a small unpack
// kernel applies the equalized-LR scale and re-lays it into the parameter's own layout.
#include "tc_common.cuh"

namespace pg {
namespace tc {

struct WgradTcParams {
  int N, H, W, Cin, Cout;      // physical channel counts of x (per tap) and of ONE dy tile (<= 256)
  int Cout_total;              // channels of dy; blockIdx.z = dy tile (wide layers: 512 = 2 x 256)
  int taps, flat;              // flat: taps address channel blocks of x ([N,1,1,taps*Cin]), no shift
  int group_stride;            // TMEM columns reserved per tap group (>= Cout)
  int bw, bh, bn, tiles_w, tiles_h, num_tiles;
  int atomM, atomN;            // channels per swizzle atom on the x / dy side (64 or 32)
  int n_atoms_m, n_atoms_n;    // atoms per 128-row x slot / per dy tile
  int atoms_per_tap;           // Cin / atomM
  int total_groups, gpp;       // tap groups overall / per pass
  int x_slots;                 // pipeline depth of the x ring
  int atom_m_bytes, atom_n_bytes;
  int tmem_cols;
  float *dwp;                  // [taps][Cout][Cin] fp32 partial-sum workspace (zeroed)
};

constexpr int kWgThreads = 256;

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slot_bytes = (uint32_t)(p.n_atoms_m * p.atom_m_bytes);   // 32 KB
  const uint32_t dy_bytes = (uint32_t)(p.n_atoms_n * p.atom_n_bytes);
  const uint32_t smem_x0 = base;
  const uint32_t smem_dy0 = base + (uint32_t)p.x_slots * slot_bytes;
  const uint32_t bar_base = smem_dy0 + 2u * dy_bytes;
  auto xfull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto xempty = [&](int s) { return bar_base + 8u * (uint32_t)(p.x_slots + s); };
  auto dyfull = [&](int d) { return bar_base + 8u * (uint32_t)(2 * p.x_slots + d); };
  auto dyempty = [&](int d) { return bar_base + 8u * (uint32_t)(2 * p.x_slots + 2 + d); };
  const uint32_t done_bar = bar_base + 8u * (uint32_t)(2 * p.x_slots + 4);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(
      smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pass = blockIdx.y;
  const int co_off = (int)blockIdx.z * p.Cout;
  const int g0 = pass * p.gpp;
  const int g1 = min(g0 + p.gpp, p.total_groups);
  const int ngroups = g1 - g0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.x_slots; ++s) {
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
        const int tn = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        // dy tile (B operand) for this pixel tile
        mbar_wait(dyempty(ds), dphase ^ 1u);
        mbar_expect_tx(dyfull(ds), dy_bytes);
        for (int a = 0; a < p.n_atoms_n; ++a)
          tma_load_4d(smem_dy0 + (uint32_t)ds * dy_bytes + (uint32_t)(a * p.atom_n_bytes), &tmap_dy,
                      dyfull(ds), co_off + a * p.atomN, w0, h0, n0);
        if (++ds == 2) {
          ds = 0;
          dphase ^= 1u;
        }
        // one x slot (128 (tap,ci) rows) per tap group
        for (int g = g0; g < g1; ++g) {
          mbar_wait(xempty(xs), xphase ^ 1u);
          int nvalid = 0;
          for (int a = 0; a < p.n_atoms_m; ++a) {
            const int tap = (g * p.n_atoms_m + a) / p.atoms_per_tap;
            if (tap < p.taps) ++nvalid;
          }
          mbar_expect_tx(xfull(xs), (uint32_t)(nvalid * p.atom_m_bytes));
          for (int a = 0; a < p.n_atoms_m; ++a) {
            const int A = g * p.n_atoms_m + a;
            const int tap = A / p.atoms_per_tap;
            if (tap >= p.taps) continue;        // dummy rows: smem left as is, lanes never read
            const int coff = (A - tap * p.atoms_per_tap) * p.atomM;
            const uint32_t dst = smem_x0 + (uint32_t)xs * slot_bytes + (uint32_t)(a * p.atom_m_bytes);
            if (p.flat)
              tma_load_4d(dst, &tmap_x, xfull(xs), tap * p.Cin + coff, w0, h0, n0);
            else
              tma_load_4d(dst, &tmap_x, xfull(xs), coff, w0 + tap % 3 - 1, h0 + tap / 3 - 1, n0);
          }
          if (++xs == p.x_slots) {
            xs = 0;
            xphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop, tcgen05 ops issued by the elected lane, descriptors = constant high
    // word + (start>>4) low word advanced by adds.
    const uint32_t idesc = make_idesc_bf16(128, p.Cout, 1, 1);   // both operands MN-major
    const uint32_t rowA = (uint32_t)p.atomM * 2u, rowB = (uint32_t)p.atomN * 2u;
    const uint32_t layA = rowA == 128 ? 2u : 4u, layB = rowB == 128 ? 2u : 4u;
    const uint32_t hiA = (((8u * rowA) >> 4) & 0x3FFFu) | (1u << 14) | (layA << 29);
    const uint32_t hiB = (((8u * rowB) >> 4) & 0x3FFFu) | (1u << 14) | (layB << 29);
    const uint32_t lboA = (((uint32_t)p.atom_m_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t lboB = (((uint32_t)p.atom_n_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t stepA = (16u * rowA) >> 4, stepB = (16u * rowB) >> 4;   // one K16 step
    int xs = 0, ds = 0;
    uint32_t xphase = 0, dphase = 0;
    uint32_t first = 1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(dyfull(ds), dphase);
      tc_fence_after();
      const uint32_t b_lo = ((smem_dy0 + (uint32_t)ds * dy_bytes) >> 4) | lboB;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(xfull(xs), xphase);
        tc_fence_after();
        const uint32_t a_lo = ((smem_x0 + (uint32_t)xs * slot_bytes) >> 4) | lboA;
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * p.group_stride);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 128 pixels = 8 x K16
            const uint64_t ad = ((uint64_t)hiA << 32) | (uint64_t)(a_lo + (uint32_t)k * stepA);
            const uint64_t bd = ((uint64_t)hiB << 32) | (uint64_t)(b_lo + (uint32_t)k * stepB);
            umma_bf16(d_tmem, ad, bd, idesc, k ? 1u : (first ^ 1u));
          }
          umma_commit(xempty(xs));
          if (g == ngroups - 1) umma_commit(dyempty(ds));
        }
        __syncwarp();
        if (++xs == p.x_slots) {
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
    const bool has_work = (int)blockIdx.x < p.num_tiles;
    if (has_work) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int m = q * 32 + lane;
      for (int g = 0; g < ngroups; ++g) {
        const int A = (g0 + g) * p.n_atoms_m + m / p.atomM;
        const int tap = A / p.atoms_per_tap;
        const int ci = (A - tap * p.atoms_per_tap) * p.atomM + m % p.atomM;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.group_stride);
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + (uint32_t)c0, v);   // warp-collective: all lanes participate
          tmem_ld_wait();
          if (tap < p.taps) {
            float *dst = p.dwp + ((size_t)tap * p.Cout_total + co_off + c0) * p.Cin + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * p.Cin, __uint_as_float(v[j]));
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

// dw[param layout, logical dims] = scale * dwp[tap][co][ci] (physical, possibly padded dims)
__global__ void __launch_bounds__(256)
wgrad_unpack_kernel(const float *__restrict__ dwp, float *__restrict__ dw, int Cin, int Cout,
                    int Cin_p, int Cout_p, int taps, float scale, int swap_io, int flip,
                    int accumulate) {
  pg::grid_dep_sync();
  const int total = Cin * Cout * taps;
  const int d1 = swap_io ? Cout : Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // i indexes the destination (coalesced writes): [d0][d1][tap']
    const int st = i % taps;
    const int i1 = (i / taps) % d1;
    const int i0 = i / (taps * d1);
    const int co = swap_io ? i1 : i0, ci = swap_io ? i0 : i1;
    const int tap = flip ? (taps - 1 - st) : st;
    const float v = scale * dwp[((size_t)tap * Cout_p + co) * Cin_p + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// deferred epilogue of every pending weight gradient in one launch (blockIdx.y = entry):
// dw += scale * ws, then ws = 0 for the next accumulation round.  Threads walk the WORKSPACE
// (coalesced read + reset; walking the gradient instead makes consecutive lanes gather at a
// 64 KB stride, which camps on one L2 slice: measured 10x slower) and update dw with a plain
// read-modify-write: the caller never puts two entries with the same dw into one launch.
__global__ void __launch_bounds__(256)
wgrad_unpack_multi_kernel(const PgUnpackEntry *__restrict__ table) {
  pg::grid_dep_sync();
  // One thread per (co, ci) pair and ALL its taps: the workspace [tap][co][ci] is read with
  // coalesced 128-byte rows (consecutive ci), the parameter gradient [d0][d1][taps] receives the
  // taps of one pair as one contiguous 36- or 64-byte run.  (One thread per workspace element
  // scattered 4-byte read-modify-writes `taps` floats apart: 54 us for the 12 trunk layers of D.)
  const PgUnpackEntry e = table[blockIdx.y];
  const int pairs = e.Cin_p * e.Cout_p;
  const int d1 = e.swap_io ? e.Cout : e.Cin;
  float *__restrict__ ws = e.ws;
  float *__restrict__ dw = e.dw;
  constexpr int kMaxTaps = 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += gridDim.x * blockDim.x) {
    const int ci = i % e.Cin_p, co = i / e.Cin_p;
    float v[kMaxTaps];                 // v[t] = the workspace tap that lands on parameter tap t
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t)
      if (t < e.taps) {
        const size_t src = (size_t)(e.flip ? (e.taps - 1 - t) : t) * pairs + i;
        v[t] = ws[src];
        ws[src] = 0.f;
      }
    if (ci < e.Cin && co < e.Cout) {
      const int i0 = e.swap_io ? ci : co, i1 = e.swap_io ? co : ci;
      float *dst = dw + ((size_t)i0 * d1 + i1) * e.taps;
      float old[kMaxTaps];
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t)
        if (t < e.taps) old[t] = dst[t];
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t)
        if (t < e.taps) dst[t] = fmaf(e.scale, v[t], old[t]);
    }
  }
}

}  // namespace tc
}  // namespace pg

using namespace pg;

extern "C" int pg_wgrad_unpack_multi(const PgUnpackEntry *table, int n, void *stream) {
  PG_CHECK_ARG(table && n > 0 && n <= 65535, "pg_wgrad_unpack_multi: bad table");   // taps <= 16 per entry
  // grid-stride over the (co, ci) pairs of each entry (1 K pairs for 32 x 32, 262 K for 512 x 512)
  int gx = 8 * sm_count() / n;
  if (gx < 16) gx = 16;
  if (gx > 1024) gx = 1024;
  dim3 grid((unsigned)gx, (unsigned)n);
  pg::launcher(tc::wgrad_unpack_multi_kernel, grid, 256, 0, (cudaStream_t)stream)(table);
  PG_CHECK_LAUNCH("pg_wgrad_unpack_multi");
}

static int pow2_ge32(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

extern "C" int pg_conv_wgrad_tc(const void *x, const void *dy, float *dw, float *workspace, int N,
                                int H, int W, int Cin, int Cout, int Cin_log, int Cout_log,
                                int taps, int flat, float scale, int swap_io, int flip,
                                int accumulate, void *stream) {
  PG_CHECK_ARG(x && dy && dw && workspace, "pg_conv_wgrad_tc: null pointer");
  PG_CHECK_ARG(taps >= 1 && (flat || taps == 9), "pg_conv_wgrad_tc: 3x3 mode needs taps == 9");
  PG_CHECK_ARG(N > 0 && H > 0 && W > 0, "pg_conv_wgrad_tc: bad dims");
  PG_CHECK_ARG(!flat || (H == 1 && W == 1), "pg_conv_wgrad_tc: flat mode needs H == W == 1");
  PG_CHECK_ARG(Cin % 32 == 0 && Cin >= 32, "pg_conv_wgrad_tc: Cin %% 32 != 0 (%d)", Cin);
  PG_CHECK_ARG(Cout % 32 == 0 && Cout >= 32 && (Cout <= 256 || (Cout % 256 == 0 && Cout <= 1024)),
               "pg_conv_wgrad_tc: Cout must be a multiple of 32 in [32,256] or of 256 up to 1024 (%d)",
               Cout);
  const int Cout_total = Cout;
  if (Cout > 256) Cout = 256;            // dy tiles of 256 channels over blockIdx.z
  const int co_tiles = Cout_total / Cout;
  PG_CHECK_ARG(Cin_log <= Cin && Cout_log <= Cout_total && Cin_log > 0 && Cout_log > 0,
               "pg_conv_wgrad_tc: logical dims exceed physical dims");
  PG_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
               "pg_conv_wgrad_tc: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  tc::WgradTcParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps; p.flat = flat;
  p.Cout_total = Cout_total;
  p.bw = W < 16 ? W : 16;
  {
    int rem = 128 / p.bw;
    p.bh = H < rem ? H : rem;
    p.bn = 128 / (p.bw * p.bh);
  }
  PG_CHECK_ARG(p.bw * p.bh * p.bn == 128 && p.bn <= 256,
               "pg_conv_wgrad_tc: cannot tile %dx%d into 128-pixel boxes", H, W);
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  const int tiles_n = (N + p.bn - 1) / p.bn;
  p.num_tiles = p.tiles_w * p.tiles_h * tiles_n;
  p.atomM = (Cin % 64 == 0) ? 64 : 32;
  p.atomN = (Cout % 64 == 0) ? 64 : 32;
  p.n_atoms_m = 128 / p.atomM;
  p.n_atoms_n = Cout / p.atomN;
  p.atoms_per_tap = Cin / p.atomM;
  const int total_atoms = taps * p.atoms_per_tap;
  p.total_groups = (total_atoms + p.n_atoms_m - 1) / p.n_atoms_m;
  p.group_stride = (Cout & (Cout - 1)) == 0 ? Cout : ((Cout + 63) / 64) * 64;
  const int max_groups = 512 / p.group_stride;
  const int passes = (p.total_groups + max_groups - 1) / max_groups;
  p.gpp = (p.total_groups + passes - 1) / passes;
  p.atom_m_bytes = 128 * p.atomM * 2;
  p.atom_n_bytes = 128 * p.atomN * 2;
  p.tmem_cols = pow2_ge32(p.gpp * p.group_stride);
  p.dwp = workspace;
  const int slot_bytes = p.n_atoms_m * p.atom_m_bytes;
  const int dy_bytes = p.n_atoms_n * p.atom_n_bytes;
  const int misc = 1024 + 8 * (2 * 6 + 5) + 64;
  int slots = (227 * 1024 - 2 * dy_bytes - misc) / slot_bytes;
  if (slots > 6) slots = 6;
  PG_CHECK_ARG(slots >= 2, "pg_conv_wgrad_tc: not enough shared memory");
  p.x_slots = slots;
  const size_t smem = (size_t)slots * slot_bytes + 2 * dy_bytes + misc;

  CUtensorMap tx, tdy;
  {
    const uint64_t cx = flat ? (uint64_t)taps * Cin : (uint64_t)Cin;
    uint64_t dims[4] = {cx, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {cx * 2, (uint64_t)W * cx * 2, (uint64_t)H * W * cx * 2};
    uint32_t box[4] = {(uint32_t)p.atomM, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (int rc = make_tmap_bf16(&tx, x, 4, dims, str, box, p.atomM * 2, "pg_conv_wgrad_tc(x)")) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout_total, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout_total * 2, (uint64_t)W * Cout_total * 2,
                       (uint64_t)H * W * Cout_total * 2};
    uint32_t box[4] = {(uint32_t)p.atomN, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (int rc = make_tmap_bf16(&tdy, dy, 4, dims, str, box, p.atomN * 2, "pg_conv_wgrad_tc(dy)")) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::wgrad_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("pg_conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
    attr_set = true;
  }
  const bool deferred = accumulate == 2;
  if (!deferred) {
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)taps * Cin * Cout_total * sizeof(float), s);
    if (e != cudaSuccess) {
      set_error("pg_conv_wgrad_tc: cudaMemsetAsync: %s", cudaGetErrorString(e));
      return PG_ERR_CUDA;
    }
  }
  int rc2 = PG_ERR_UNSUPPORTED;
  if (!flat && taps == 9 && co_tiles == 1) {  // newer kernel generations where the shape allows
    rc2 = wgrad4_tc_launch(x, dy, workspace, N, H, W, Cin, Cout, s);
  }
  if (rc2 != PG_OK && rc2 != PG_ERR_UNSUPPORTED) return rc2;
  if (rc2 == PG_ERR_UNSUPPORTED) {
    int gx = sm_count() / (passes * co_tiles);
    if (gx > p.num_tiles) gx = p.num_tiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, passes, co_tiles);
    pg::launcher(tc::wgrad_tc_kernel, grid, tc::kWgThreads, smem, s, p.num_tiles)(tx, tdy, p);
  }
  const int total = taps * Cin_log * Cout_log;
  if (!deferred)
  pg::launcher(tc::wgrad_unpack_kernel, (total + 255) / 256, 256, 0, s)(workspace, dw, Cin_log, Cout_log, Cin,
                                                             Cout_total, taps, scale, swap_io, flip, accumulate);
  PG_CHECK_LAUNCH("pg_conv_wgrad_tc");
}
