// Micro-benchmark: TMEM -> register read throughput (tcgen05.ld) per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fsnerf_b200/csrc/common.cuh"
void fsnerf_set_error(const char*, ...) {}
using namespace fs;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr) : "memory");
}

template <int MODE>  // 0: x32 + wait each; 1: 2 x32 in flight; 2: x64 + wait each; 3: 4 x32 then wait
__global__ void bench(long long* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) { uint32_t v[32]; tmem_ld32(base + c * 32, v); tmem_ld_wait(); acc += v[0] ^ v[31]; }
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 8; c += 2) { uint32_t v[32], w[32]; tmem_ld32(base + c * 32, v); tmem_ld32(base + c * 32 + 32, w); tmem_ld_wait(); acc += v[0] ^ w[31]; }
      } else if (MODE == 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { uint32_t v[64]; tmem_ld64(base + c * 64, v); tmem_ld_wait(); acc += v[0] ^ v[63]; }
      } else {
#pragma unroll
        for (int c = 0; c < 8; c += 4) { uint32_t v[32], w[32], x[32], y[32]; tmem_ld32(base + c * 32, v); tmem_ld32(base + c * 32 + 32, w); tmem_ld32(base + c * 32 + 64, x); tmem_ld32(base + c * 32 + 96, y); tmem_ld_wait(); acc += v[0] ^ w[31] ^ x[1] ^ y[2]; }
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 8192 * 8);
  const int iters = 200;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    for (int nw : {4, 8}) {
      for (int mode = 0; mode < 4; ++mode) {
        int threads = nw * 32;
        int grid = 148 * ctas;
        if (mode == 0) bench<0><<<grid, threads>>>(d, iters, nw);
        if (mode == 1) bench<1><<<grid, threads>>>(d, iters, nw);
        if (mode == 2) bench<2><<<grid, threads>>>(d, iters, nw);
        if (mode == 3) bench<3><<<grid, threads>>>(d, iters, nw);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        // per iteration each warp reads 256 cols x 32 lanes x 4 B = 32 KB
        double bytes_per_cta = (double)iters * nw * 32768.0;
        printf("ctas/SM %d warps %d mode %d: %lld cycles, %.1f cycles per 128x256 fp32 tile-read (4 warps), %.1f B/clk/CTA  %s\n",
               ctas, nw, mode, h[0], (double)h[0] / iters / (nw / 4.0), bytes_per_cta / h[0], cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
