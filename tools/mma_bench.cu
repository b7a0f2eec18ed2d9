// Micro-benchmark: tcgen05.mma issue/execute rate per SM for the shapes the fused MLP uses,
// alone and with the competing traffic of the real kernel (TMEM epilogue loads, bulk copies
// into the shared-memory weight ring).
//   mode bit 0: A operand from TMEM (TS) instead of shared memory (SS)
//   mode bit 1: 4 extra warps stream tcgen05.ld over the other 256 TMEM columns
//   mode bit 2: 2 extra warps keep 32 KB bulk copies L2 -> smem in flight
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fsnerf_b200/csrc/common.cuh"
void fsnerf_set_error(const char*, ...) {}
using namespace fs;

constexpr int kB = 32768;

// converged-code variant: all 32 lanes execute, operands stay in uniform registers, the
// instruction itself is guarded by a per-lane predicate (lane 0)
__device__ __forceinline__ void umma_ts_pred(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                             uint32_t issue) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc),
               "r"(idesc), "r"(acc), "r"(issue) : "memory");
}

template <int N>
__global__ void __launch_bounds__(256, 1) bench(const uint8_t* src, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // layout: A tile 16 KB | B tile 32 KB | 4 x 32 KB copy ring | barriers
  const uint32_t a_tile = sbase, b_tile = sbase + 16384, ring = sbase + 49152;
  const uint32_t bars = ring + 4 * kB;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
    stop = 0;
  }
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    t0 = clock64();
    if (mode & 8) {
      const uint32_t issue = lane == 0;
      const int delay = (mode >> 8) * 50;  // idle cycles between groups of 4 MMAs
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts_pred(tm, tm + 256 + 8 * k, umma_desc_sw128(b_tile + k * 32, 16, 1024), idesc, 1, issue);
        if (delay) { long long tw = clock64(); while (clock64() - tw < delay) {} }
        if (it == 0 && iters <= 4) { long long ti = clock64(); if (lane == 0) out[2000 + blockIdx.x] = ti - t0; }
      }
      if (lane == 0) umma_commit(bars);
    } else if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode & 1)
            umma_bf16_ts(tm, tm + 256 + 8 * k, umma_desc_sw128(b_tile + k * 32, 16, 1024), idesc, 1);
          else
            umma_bf16_ss(tm, umma_desc_sw128(a_tile + k * 32, 16, 1024), umma_desc_sw128(b_tile + k * 32, 16, 1024),
                         idesc, 1);
        }
        if (it == 0 && iters <= 4) out[2000 + blockIdx.x] = clock64() - t0;
      }
      umma_commit(bars);
    }
    __syncwarp();
    mbar_wait(bars, 0);
    t1 = clock64();
    stop = 1;
  } else if (warp >= 4 && (mode & 2)) {
    const uint32_t base = tm + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    while (!__shfl_sync(0xffffffffu, (int)stop, 0)) {
      uint32_t v[32];
      tmem_ld32(base + 64, v); tmem_ld_wait(); acc += v[3];
      tmem_ld32(base + 96, v); tmem_ld_wait(); acc += v[5];
    }
    if (acc == 0x1234567) out[1000] = acc;
  } else if ((warp == 2 || warp == 3) && (mode & 4)) {
    const int me = warp - 2;
    uint32_t n = 0;
    while (!__shfl_sync(0xffffffffu, (int)stop, 0)) {
      const uint32_t s = me * 2 + (n & 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(bars + 8 * (1 + s), kB);
        bulk_g2s(ring + s * kB, src + (size_t)((n * 2 + me + blockIdx.x) % 32) * kB, kB, bars + 8 * (1 + s));
        mbar_wait(bars + 8 * (1 + s), (n >> 1) & 1);
      }
      __syncwarp();
      ++n;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// the fused kernel's per-chunk issue pattern, bisected by `what` bits:
//   1: tcgen05.fence::after_thread_sync per group   2: probes (test_wait x2) inside the group
//   4: tcgen05.commit per group                     8: B operand rotates over 4 ring stages
__global__ void __launch_bounds__(384, 1) pattern(int iters, int what, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t ring = sbase + 49152, bars = ring + 4 * kB;
  __shared__ uint32_t slot;
  __shared__ int slot2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    slot2 = 0;
    for (int s = 0; s < 16; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < (49152 + 4 * kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    const uint32_t issue = lane == 0;
    long long t0 = clock64();
    uint32_t ok0 = 0, ok1 = 0, sink = 0;
    for (int it = 0; it < iters; ++it) {
      if (what & 1) tc_fence_after();
      const uint32_t b_tile = (what & 8) ? ring + (it & 3) * kB : sbase + 16384;
      uint64_t ad[4], bd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ad[k] = tm + 256 + 64 * (it & 3) + 32 * (k >> 1) + 8 * (k & 1);
        bd[k] = umma_desc_sw128(b_tile + k * 32, 16, 1024);
      }
      if (what & 2) {
        umma_bf16_ts_conv(tm, (uint32_t)ad[0], bd[0], idesc, 1, issue);
        umma2_probe<true>(tm, ad[1], ad[2], bd[1], bd[2], idesc, issue, bars + 64, 0, bars + 72, 0, ok0, ok1);
        umma_bf16_ts_conv(tm, (uint32_t)ad[3], bd[3], idesc, 1, issue);
        sink += ok0 + ok1;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts_conv(tm, (uint32_t)ad[k], bd[k], idesc, 1, issue);
      }
      if (what & 4) umma_commit_conv(bars + 8 * (1 + (it & 3)), issue);
    }
    umma_commit_conv(bars, issue);
    mbar_wait(bars, 0);
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = sink; }
    *(volatile int*)&slot2 = 1;
  } else if (warp >= 4 && warp < 12 && (what & (16 | 32 | 64))) {
    // epilogue-like traffic in the region the MMAs read A from (columns 256..511)
    const uint32_t base = tm + 256 + ((uint32_t)((warp & 3) * 32) << 16) + 32 * ((warp >> 2) & 1);
    uint32_t acc = 0, n = 0;
    while (!__shfl_sync(0xffffffffu, *(volatile int*)&slot2, 0)) {
      const uint32_t col = base + 64 * (n & 3);
      uint32_t v[32], w[16];
      if (what & (16 | 32)) { tmem_ld32(col, v); tmem_ld_wait(); } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = n + i;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]) + 1.0f, __uint_as_float(v[2 * i + 1]) + 1.0f);
      if (what & (16 | 64)) { tmem_st16(col, w); tmem_st_wait(); } else acc += w[3] ^ w[7];
      ++n;
    }
    if (acc == 0x1234567) out[1000] = acc;
  }
  __syncthreads();
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// the fused kernel's issue loop, verbatim structure (carried state, waits guarded by probe
// results, commits), running alone on the SM: what & 1 -> N=128 MMAs
__global__ void __launch_bounds__(480, 1) issue_loop(int iters, int what, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t ring = sbase, bars = ring + 5 * kB;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 24; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < (5 * kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 12) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 12) {
    const uint32_t issue = lane == 0;
    const uint32_t idesc = (what & 1) ? umma_idesc_bf16(128, 128, 0, 0) : umma_idesc_bf16(128, 256, 0, 0);
    // pre-complete phase 0 of the "operand ready" barriers so that probes of parity 0 succeed
    if (lane == 0) { for (int s = 0; s < 10; ++s) mbar_arrive(bars + 8 * s); }
    __syncwarp();
    uint32_t stage = 0, ok_a = 0, ok_w = 0;
    uint32_t flags = 0, a0 = tm + 256, b0 = umma_desc_lo(ring), abar = bars, wbar = bars + 40, ebar = bars + 80;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const bool first = (it & 3) == 0;
      const uint32_t d_col = tm + (flags & 0x100);
      if (!(what & 16)) {
        if (!ok_a) mbar_wait(abar, 0);
        if (!ok_w) mbar_wait(wbar, 0);
      }
      if (!(what & 8)) tc_fence_after();
      umma_bf16_ts_conv(d_col, a0, umma_desc_from_lo(b0), idesc, first ? 0u : 1u, issue);
      const uint32_t ebar_cur = ebar;
      if (++stage == 5) stage = 0;
      const uint32_t c = (it + 1) & 3;
      const uint32_t nflags = (((it + 1) >> 2) & 1) ? 0x100u : 0u;
      const uint32_t na0 = tm + (256 - (nflags & 0x100)) + 64 * c;
      const uint32_t nb0 = umma_desc_lo(ring + stage * kB);
      const uint32_t nabar = bars + 8 * c, nwbar = bars + 40 + 8 * stage, nebar = bars + 80 + 8 * stage;
      uint32_t pa, pw;
      if (what & 4) {
        umma_bf16_ts_conv(d_col, a0 + 8, umma_desc_from_lo(b0 + 2), idesc, 1u, issue);
        umma_bf16_ts_conv(d_col, a0 + 32, umma_desc_from_lo(b0 + 4), idesc, 1u, issue);
        pa = pw = 1;
      } else {
        umma2_probe<true>(d_col, a0 + 8, a0 + 32, umma_desc_from_lo(b0 + 2), umma_desc_from_lo(b0 + 4), idesc, issue,
                          nabar, 0, nwbar, 0, pa, pw);
      }
      umma_bf16_ts_conv(d_col, a0 + 40, umma_desc_from_lo(b0 + 6), idesc, 1u, issue);
      if (!(what & 2)) {
        umma_commit_conv(ebar_cur, issue);
        if (c == 0) umma_commit_conv(bars + 160 + 8 * ((it >> 2) & 1), issue);
      }
      flags = nflags; a0 = na0; b0 = nb0; abar = nabar; wbar = nwbar; ebar = nebar;
      ok_a = pa; ok_w = pw;
    }
    umma_commit_conv(bars + 184, issue);
    mbar_wait(bars + 184, 0);
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before(); __syncthreads();
  if (warp == 12) tmem_dealloc(tm, 512);
}

int main() {
  uint8_t* d; cudaMalloc(&d, 4 << 20); cudaMemset(d, 0, 4 << 20);
  long long* dout; cudaMalloc(&dout, 16384);
  const int smem = 49152 + 4 * kB + 256;
  cudaFuncSetAttribute(bench<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // issue cost of 4 MMAs into an empty queue + total latency of 4 MMAs
  for (int mode : {1, 9}) {
    bench<256><<<1, 256, smem>>>(d, 1, mode, dout);
    cudaDeviceSynchronize();
    long long h1, h2; cudaMemcpy(&h1, dout, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h2, dout + 2000, 8, cudaMemcpyDeviceToHost);
    printf("N=256 TS %s: 4 MMAs issued in %lld cycles, complete (commit observed) after %lld cycles\n",
           (mode & 8) ? "converged+predicated" : "lane-0 branch", h2, h1);
  }
  cudaFuncSetAttribute(issue_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * kB + 512);
  for (int what : {0, 1, 3, 5, 9, 17, 31}) {
    issue_loop<<<148, 480, 5 * kB + 512>>>(2000, what, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long h1; cudaMemcpy(&h1, dout, 8, cudaMemcpyDeviceToHost);
    printf("issue_loop N=%d%s%s%s%s: %.0f cycles per chunk (4 MMAs)  %s\n", (what & 1) ? 128 : 256, (what & 2) ? " -commit" : "",
           (what & 4) ? " -probes" : "", (what & 8) ? " -fence" : "", (what & 16) ? " -okchecks" : "", (double)h1 / 2000, cudaGetErrorString(e));
  }
  cudaFuncSetAttribute(pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int what : {15, 15 + 16, 15 + 32, 15 + 64}) {
    pattern<<<148, 384, smem>>>(2000, what, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long h1; cudaMemcpy(&h1, dout, 8, cudaMemcpyDeviceToHost);
    printf("pattern[%3d] %s%s%s%s: %.0f cycles per 4-MMA group  %s\n", what, (what & 1) ? "fence " : "", (what & 2) ? "probes " : "",
           (what & 4) ? "commit " : "", (what & 8) ? "ring " : "", (double)h1 / 2000, cudaGetErrorString(e));
  }
  for (int dl = 0; dl <= 0; ++dl) {
    bench<256><<<1, 256, smem>>>(d, 500, 9 | (dl << 8), dout);
    cudaDeviceSynchronize();
    long long h1; cudaMemcpy(&h1, dout, 8, cudaMemcpyDeviceToHost);
    printf("4 MMAs (552 cycles of work) + %3d idle issue cycles per group: %.0f cycles per group\n", dl * 50, (double)h1 / 500);
  }
  const int iters = 2000;
  for (int grid : {1}) {
    for (int n : {256}) {
      for (int mode : {1, 9}) {
        if (n == 128) bench<128><<<grid, 256, smem>>>(d, iters, mode, dout);
        else bench<256><<<grid, 256, smem>>>(d, iters, mode, dout);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, dout, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        double per = (double)mx / (iters * 4.0);
        printf("grid %3d N=%3d %s%s%s: %.1f cycles per MMA (ideal %d), %.0f MAC/clk/SM  %s\n", grid, n,
               (mode & 1) ? "TS" : "SS", (mode & 2) ? "+tmem_ld" : "", (mode & 4) ? "+bulk" : "", per, n / 2,
               128.0 * n * 16 / per, cudaGetErrorString(e));
        fflush(stdout);
      }
    }
  }
  return 0;
}
