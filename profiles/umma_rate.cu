// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, bf16, both
// operands from shared memory) as a function of N and of how many independent TMEM
// accumulators consecutive instructions alternate between.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate profiles/umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../progressive-gan-pytorch_b200/csrc/tc_common.cuh"

namespace pg { void set_error(const char *, ...) {} }
using namespace pg::tc;

// mode 0: every MMA reads the same A/B tiles; mode 1: A start walks over 9 row-shifted windows
// (like the conv), B walks over distinct tiles
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int N, int nacc, int n_mma, int reps, int mode, int sbo_a, long long *cycles, long long *nanos) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;                     // 32 KB region
  const uint32_t smem_b = base + 32768u;            // up to 9 x 256 x 128 B
  const uint32_t bar = base + 32768u + 4u * 32768u;
  const uint32_t slot = bar + 8u;
  volatile uint32_t *slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  // fill smem with something finite
  for (uint32_t i = threadIdx.x; i < (32768u + 4u * 32768u) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t hi_a = (((uint32_t)sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t hi_b = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    uint32_t phase = 0;
    long long t0 = 0, t1 = 0, g0 = 0, g1 = 0;
    for (int r = 0; r < reps + 1; ++r) {
      if (r == 1) {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      }
      if (elect_one_sync()) {
        const uint32_t a0 = (smem_a >> 4) | (1u << 16), b0 = (smem_b >> 4) | (1u << 16);
        for (int i = 0; i < n_mma; i += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = j & 3;
            const int tap = mode ? (j >> 2) : 0;                 // compile-time after unrolling
            const uint32_t a_off = (uint32_t)(mode ? ((tap / 3) * 10 + tap % 3) * 128 : 0) + k * 32u;
            const uint32_t b_off = (uint32_t)(mode ? tap * 32768 : 0) + k * 32u;
            const uint64_t ad = ((uint64_t)hi_a << 32) | (uint64_t)(a0 + (a_off >> 4));
            const uint64_t bd = ((uint64_t)hi_b << 32) | (uint64_t)(b0 + (b_off >> 4));
            const int acc = (nacc > 0) ? (j & (nacc - 1)) : ((j >> 2) & (-nacc - 1));
            umma_bf16(tmem_base + (uint32_t)(acc * N), ad, bd, idesc, 1u);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) {
      cycles[blockIdx.x] = t1 - t0;
      nanos[blockIdx.x] = g1 - g0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// Emulates conv4's tile structure: groups of `gsz` MMAs alternate between two accumulators, one
// commit per group (tfull), the issuer waits for the drain of the stage it reuses (tempty).
// drain = 1: warps 4..11 read the finished accumulator with tcgen05.ld (like the epilogue) and
// arrive on tempty; drain = 0: a single thread arrives immediately.
__global__ void __launch_bounds__(384, 1)
umma_tile_kernel(int N, int gsz, int ngroups, int drain, long long *cycles, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base, smem_b = base + 32768u;
  const uint32_t bar = base + 32768u + 4u * 32768u;     // tfull[2], tempty[2]
  const uint32_t slot = bar + 64u;
  volatile uint32_t *slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (32768u + 4u * 32768u) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  const int n_arrive = drain ? 256 : 1;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(bar + 8, 1);
    mbar_init(bar + 16, n_arrive); mbar_init(bar + 24, n_arrive);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t hi_a = (((uint32_t)1280 >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t hi_b = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t a0 = (smem_a >> 4) | (1u << 16), b0 = (smem_b >> 4) | (1u << 16);
    int acc = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    if (variant & 4) {
      // ONE election around the whole persistent loop: the issuing thread never re-converges
      if (elect_one_sync()) {
        for (int g = 0; g < ngroups; ++g) {
          if (variant & 2) {
            mbar_wait(bar + 16 + 8 * acc, ph ^ 1u);
            tc_fence_after();
          }
          for (int i = 0; i < gsz; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int tap = (i >> 2) % 9;
              const uint32_t a_off = (uint32_t)((tap / 3) * 10 + tap % 3) * 128u + k * 32u;
              const uint32_t b_off = (uint32_t)(tap & 3) * 32768u + k * 32u;
              umma_bf16(tmem_base + (uint32_t)(acc * N), ((uint64_t)hi_a << 32) | (uint64_t)(a0 + (a_off >> 4)),
                        ((uint64_t)hi_b << 32) | (uint64_t)(b0 + (b_off >> 4)), idesc, 1u);
            }
          }
          if (variant & 1) umma_commit(bar + 8 * acc);
          if (++acc == 2) { acc = 0; ph ^= 1u; }
        }
      }
      __syncwarp();
    } else
    for (int g = 0; g < ngroups; ++g) {
      if (variant & 2) {
        mbar_wait(bar + 16 + 8 * acc, ph ^ 1u);
        tc_fence_after();
      }
      if (elect_one_sync()) {
        for (int i = 0; i < gsz; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int tap = (i >> 2) % 9;
            const uint32_t a_off = (uint32_t)((tap / 3) * 10 + tap % 3) * 128u + k * 32u;
            const uint32_t b_off = (uint32_t)(tap & 3) * 32768u + k * 32u;
            umma_bf16(tmem_base + (uint32_t)(acc * N), ((uint64_t)hi_a << 32) | (uint64_t)(a0 + (a_off >> 4)),
                      ((uint64_t)hi_b << 32) | (uint64_t)(b0 + (b_off >> 4)), idesc, 1u);
          }
        }
        if (variant & 1) umma_commit(bar + 8 * acc);
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; ph ^= 1u; }
    }
    // wait for the last two groups
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  } else if ((variant & 1) && (warp >= 4 || (warp == 0 && !drain))) {
    if (drain || lane == 0) {
      int acc = 0; uint32_t ph = 0;
      const int q = warp & 3;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(bar + 8 * acc, ph);
        tc_fence_after();
        if (drain) {
          const int half = (warp - 4) >> 2;
          uint32_t v[32];
          for (int c = 0; c < N / 2; c += 32) {
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N + half * (N / 2) + c), v);
            tmem_ld_wait();
          }
          if (v[0] == 0x12345678u) cycles[200] = 1;     // keep the loads alive
        }
        tc_fence_before();
        mbar_arrive(bar + 16 + 8 * acc);
        if (++acc == 2) { acc = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long *cyc, *ns;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&ns, 148 * 8);
  const size_t smem = 1024 + 32768 + 4 * 32768 + 64;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n_mma = 1024, reps = 20;
  printf("%-6s %-5s %-5s %-6s %-6s %10s %10s %10s\n", "grid", "N", "nacc", "mode", "sbo", "cyc/mma", "ns/mma", "TF/s/chip");
  for (int grid : {1, 148})
    for (int mode : {0, 1})
      for (int N : {32, 64, 128, 256})
        for (int nacc : {1, 2, 4, -2}) {
          if ((nacc < 0 ? -nacc : nacc) * N > 512) continue;
          const int sbo = mode ? 1280 : 1024;
          umma_rate_kernel<<<grid, 128, smem>>>(N, nacc, n_mma, reps, mode, sbo, cyc, ns);
          cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return 1; }
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long hc[148], hn[148];
          cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
          cudaMemcpy(hn, ns, grid * 8, cudaMemcpyDeviceToHost);
          double c = 0, n = 0;
          for (int i = 0; i < grid; ++i) { c += hc[i]; n += hn[i]; }
          c /= grid; n /= grid;
          const double per = c / (double)(n_mma * reps), pern = n / (double)(n_mma * reps);
          printf("%-6d %-5d %-5d %-6d %-6d %10.1f %10.1f %10.1f\n", grid, N, nacc, mode, sbo, per, pern,
                 2.0 * 128 * N * 16 / pern * 148 / 1e3);
        }
  {
    long long *cyc2; cudaMalloc(&cyc2, 256 * 8);
    const size_t smem2 = 1024 + 32768 + 4 * 32768 + 256;
    cudaFuncSetAttribute(umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    printf("tile structure: N gsz drain -> cycles per MMA\n");
    for (int N : {64, 128})
      for (int gsz : {36, 72})
        for (int variant : {3, 7}) {
          const int ngroups = 64, drain = 0;
          umma_tile_kernel<<<148, 384, smem2>>>(N, gsz, ngroups, drain, cyc2, variant);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long hc[148]; cudaMemcpy(hc, cyc2, 148 * 8, cudaMemcpyDeviceToHost);
          double c = 0; for (int i = 0; i < 148; ++i) c += hc[i]; c /= 148;
          printf("%-4d %-4d variant %d (1 commit, 2 wait) %8.1f\n", N, gsz, variant, c / (double)(gsz * ngroups));
        }
  }
  return 0;
}
