// Micro-benchmark: L2 -> shared memory bulk-copy (cp.async.bulk) throughput when every SM
// streams the same ~1.2 MB weight image over and over (the fused MLP's operand traffic).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fsnerf_b200/csrc/common.cuh"
void fsnerf_set_error(const char*, ...) {}
using namespace fs;

template <int kStages, int kBytes, int kProd>
__global__ void __launch_bounds__(32 * kProd, 1) bench(const uint8_t* src, int n_blocks, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int w = threadIdx.x >> 5;
  const uint32_t sbase = smem_u32(smem) + w * kStages * kBytes;
  const uint32_t bars = smem_u32(smem) + kProd * kStages * kBytes + w * kStages * 8;
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0) {
    // keep kStages copies in flight
    int issued = 0, done = 0;
    const int total = iters;
    for (; issued < kStages && issued < total; ++issued) {
      mbar_arrive_expect_tx(bars + 8 * (issued % kStages), kBytes);
      bulk_g2s(sbase + (issued % kStages) * kBytes, src + (size_t)((issued + blockIdx.x * 7) % n_blocks) * kBytes, kBytes,
               bars + 8 * (issued % kStages));
    }
    while (done < total) {
      const int s = done % kStages;
      mbar_wait(bars + 8 * s, (done / kStages) & 1);
      ++done;
      if (issued < total) {
        mbar_arrive_expect_tx(bars + 8 * s, kBytes);
        bulk_g2s(sbase + s * kBytes, src + (size_t)((issued + blockIdx.x * 7) % n_blocks) * kBytes, kBytes, bars + 8 * s);
        ++issued;
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int kStages, int kBytes, int kProd = 1>
void run(const uint8_t* d, long long* dout, int grid) {
  const int smem = kProd * kStages * kBytes + 256;
  cudaFuncSetAttribute(bench<kStages, kBytes, kProd>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const int n_blocks = (1200 * 1024) / kBytes;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<kStages, kBytes, kProd><<<grid, 32 * kProd, smem>>>(d, n_blocks, 50, dout);
  cudaEventRecord(e0);
  bench<kStages, kBytes, kProd><<<grid, 32 * kProd, smem>>>(d, n_blocks, iters, dout);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, dout, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < (grid < 148 ? grid : 148); ++i) mx = h[i] > mx ? h[i] : mx;
  double bytes = (double)iters * kBytes * kProd;
  printf("grid %3d prod %d stages %d x %6d B: %.1f B/clk/SM, %.0f B/clk chip, %.2f TB/s (%.3f ms)  %s\n", grid, kProd, kStages, kBytes,
         bytes / mx, bytes * grid / mx, bytes * grid / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(e));
}

int main() {
  uint8_t* d; cudaMalloc(&d, 4 << 20); cudaMemset(d, 1, 4 << 20);
  long long* dout; cudaMalloc(&dout, 8192);
  for (int grid : {1, 148}) {
    run<4, 4096>(d, dout, grid);
    run<4, 16384>(d, dout, grid);
    run<3, 32768>(d, dout, grid);
    run<3, 65536>(d, dout, grid);
    run<1, 131072>(d, dout, grid);
    run<2, 16384, 2>(d, dout, grid);
    run<2, 16384, 4>(d, dout, grid);
    run<2, 16384, 6>(d, dout, grid);
    run<2, 32768, 3>(d, dout, grid);
    run<1, 4096, 8>(d, dout, grid);
    run<4, 4096, 8>(d, dout, grid);
  }
  return 0;
}
