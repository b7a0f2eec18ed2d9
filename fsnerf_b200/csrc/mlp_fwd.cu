// Kernels (2)+(3): frequency-masked positional encoding fused into the fully
// fused NeRF MLP forward (reference: src/core/models.py:111-143 evaluated in
// the closures of src/render/rendering.py:58-84).
//
// One CTA = one 128-sample tile at a time, 2 CTAs co-resident per SM so that
// one CTA's epilogue overlaps the other's MMAs.  Per CTA (192 threads):
//   warp 0      weight producer: streams 16 KB pre-swizzled bf16 operand blocks
//               L2 -> smem with cp.async.bulk (TMA engine) through a ring
//   warp 1      MMA issuer: tcgen05.mma (M=128, N=128, K=16, bf16 -> fp32 in
//               TMEM), commits ring slots and the accumulator to mbarriers
//   warps 2..5  one thread per sample row: ray -> position -> sin/cos encoding
//               (x freq mask) -> bf16 A tile; per layer TMEM -> regs -> bias +
//               ReLU -> bf16 -> the next layer's A tile in smem (activations
//               never leave the SM); sigma / rgb heads on CUDA cores.
// Optional stash: each A tile image is bulk-stored to HBM for the backward.
#include <stdlib.h>
#include "common.cuh"
#include "mlp_common.cuh"
#include "mlp_encode.cuh"

namespace fs {
namespace {

constexpr int kFwdThreads = 192;
constexpr int kSmemAct = 0;                                   // 4 chunks x 16 KB
constexpr int kSmemAux = kSmemAct + 4 * kChunkBytes;          // 16 KB
constexpr int kSmemRing = kSmemAux + kChunkBytes;             // kStages x 16 KB
template <int kStages> __host__ __device__ constexpr int smem_bars() { return kSmemRing + kStages * kBlockBytes; }
template <int kStages> __host__ __device__ constexpr int smem_total() { return smem_bars<kStages>() + 256; }
constexpr int kTmemCols = 256;

// fp32 biases + sigma/rgb head weights of the network being evaluated (copied
// device-to-device from the packed image's small-params block before each launch)
__constant__ float c_small[kSmallFloats];

// optional per-phase clock64 trace of CTA 0 (fsnerf_debug_set_trace), tuning aid
#define FS_TRACE(slot, g_, k_)                                                              \
  do {                                                                                      \
    if (args.trace && blockIdx.x == 0 && (slot) < 4)                                        \
      args.trace[((slot) * 16 + (g_)) * 8 + (k_)] = clock64();                              \
  } while (0)

struct FwdArgs {
  long long* trace;
  const float* params;
  const uint8_t* packed;
  int64_t n_samples;
  int samples_per_ray;
  const float* rays_o;
  const float* rays_d;
  const float* t_starts;
  const float* t_ends;
  const float* x;
  const float* dirs;
  const float* mask_pos;
  const float* mask_dir;
  int density_only;
  float* out;
  uint8_t* stash;
};

template <int kStages, int kMinBlocks>
__global__ void __launch_bounds__(kFwdThreads, kMinBlocks)
mlp_fwd_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ FwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kSmemBars = smem_bars<kStages>();
  const uint32_t bar_w_full = sbase + kSmemBars;            // [kStages]
  const uint32_t bar_w_empty = bar_w_full + 8 * kStages;    // [kStages]
  const uint32_t bar_a_ready = bar_w_empty + 8 * kStages;
  const uint32_t bar_acc_full = bar_a_ready + 8;
  const uint32_t tmem_slot = bar_acc_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 8 * (2 * kStages + 2));

  const int n_gemm = args.density_only ? prog.n_hidden : prog.n_gemm;
  int dir_after = 0;  // last hidden layer whose MMA reads the aux (position encoding) buffer
  for (int g = 0; g < prog.n_hidden; ++g)
    if (prog.layer[g].use_aux) dir_after = g;
  const int n_blocks = args.density_only ? prog.n_blocks_fwd_density : prog.n_blocks_fwd;
  const int64_t n_tiles = (args.n_samples + kTileM - 1) / kTileM;

  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("fsnerf: dynamic smem base not 1024B aligned (%u)\n", sbase);
    __trap();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_w_full + 8 * s, 1);
      mbar_init(bar_w_empty + 8 * s, 1);
    }
    mbar_init(bar_a_ready, 128);
    mbar_init(bar_acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------ weight producer
    uint32_t cnt = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int b = 0; b < n_blocks; ++b, ++cnt) {
        const uint32_t stage = cnt % kStages, phase = (cnt / kStages) & 1;
        mbar_wait(bar_w_empty + 8 * stage, phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_w_full + 8 * stage, kBlockBytes);
          bulk_g2s(sbase + kSmemRing + stage * kBlockBytes, args.packed + (size_t)b * kBlockBytes,
                   kBlockBytes, bar_w_full + 8 * stage);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    uint32_t cnt = 0, a_phase = 0;
    int titer = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int g = 0; g < n_gemm; ++g) {
        const GemmLayer& L = prog.layer[g];
        if (lane == 0) FS_TRACE(titer, g, 0);
        mbar_wait(bar_a_ready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
        if (lane == 0) FS_TRACE(titer, g, 1);
        const int nchunks = L.n_act_chunks + L.use_aux;
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t a_tile =
              sbase + ((c < L.n_act_chunks) ? (kSmemAct + c * kChunkBytes) : kSmemAux);
          for (int nh = 0; nh < L.n_halves; ++nh, ++cnt) {
            const uint32_t stage = cnt % kStages, phase = (cnt / kStages) & 1;
            mbar_wait(bar_w_full + 8 * stage, phase);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t b_tile = sbase + kSmemRing + stage * kBlockBytes;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_ss(tmem_base + nh * 128, umma_desc_sw128(a_tile + k * 32, 16, 1024),
                             umma_desc_sw128(b_tile + k * 32, 16, 1024), idesc,
                             (c > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(bar_w_empty + 8 * stage);
            }
            __syncwarp();
          }
        }
        if (lane == 0) umma_commit(bar_acc_full);
        if (lane == 0) FS_TRACE(titer, g, 2);
        __syncwarp();
      }
      ++titer;
    }
  } else {
    // ------------------------------------------------ encode + epilogue (thread = sample row)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_phase = 0;
    int titer = -1;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      ++titer;
      if (et == 0) FS_TRACE(titer, 15, 6);
      const int64_t p = tile * kTileM + row;
      const bool valid = p < args.n_samples;
      uint8_t* stash_tile = args.stash ? args.stash + (size_t)tile * prog.stash_tile_bytes : nullptr;
      float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
      if (valid) {
        if (args.x) {
#pragma unroll
          for (int a = 0; a < 3; ++a) pos[a] = args.x[p * 3 + a];
          if (args.dirs) {
#pragma unroll
            for (int a = 0; a < 3; ++a) dir[a] = args.dirs[p * 3 + a];
          }
        } else {
          const int64_t ray = p / args.samples_per_ray;
          // reference: src/render/rendering.py:79  x = o + d*(ts+te)/2
          const float tm = (args.t_starts[p] + args.t_ends[p]) / 2.0f;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            dir[a] = args.rays_d[ray * 3 + a];
            pos[a] = args.rays_o[ray * 3 + a] + dir[a] * tm;
          }
        }
      }
      if (stash_tile) {  // previous tile's image stores must have finished reading aux/act
        if (et == 0) bulk_wait_read0();
        named_bar_sync(1, 128);
      }
      encode_row(pos, prog.n_freqs_pos, prog.freq_pos, prog.pow2_freqs != 0, args.mask_pos, sbase + kSmemAux, row);
      fence_proxy_async_smem();
      mbar_arrive(bar_a_ready);
      if (et == 0) FS_TRACE(titer, 15, 7);
      if (stash_tile) {
        named_bar_sync(1, 128);
        if (et == 0) {
          bulk_s2g(stash_tile + prog.stash_aux_pos_off, sbase + kSmemAux, kChunkBytes);
          bulk_commit();
        }
      }
      float sigma = 0.f;
      for (int g = 0; g < n_gemm; ++g) {
        const GemmLayer& L = prog.layer[g];
        if (et == 0) FS_TRACE(titer, g, 3);
        mbar_wait(bar_acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (et == 0) FS_TRACE(titer, g, 4);
        if (stash_tile) {  // the previous image store must have finished reading act
          if (et == 0) bulk_wait_read0();
          named_bar_sync(1, 128);
        }
        const int ncols = L.n_halves * 128;
        const int epi = L.epi;
        float rgb_acc[3] = {0.f, 0.f, 0.f};
        // software pipeline: the TMEM load of chunk ci+1 is in flight while chunk ci is
        // converted; bias / head weights are uniform __constant__ operands (no loads)
        uint32_t v[2][32];
        tmem_ld32(tmem_row, v[0]);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c0 = ci * 32;
          if (c0 < ncols) {
            tmem_ld_wait();
            if (c0 + 32 < ncols) tmem_ld32(tmem_row + c0 + 32, v[(ci + 1) & 1]);
            const uint32_t(&vc)[32] = v[ci & 1];
            const float2* __restrict__ cb2 =
                reinterpret_cast<const float2*>(c_small + kSmallBias + g * 256 + c0);
            uint32_t w[16];
            if (epi == EPI_RELU || epi == EPI_CONN) {
              // hot path: packed fp32 bias add (FADD2) + relu fused into the bf16x2 convert
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 b = cb2[i];
                float s0, s1;
                add_f32x2(s0, s1, __uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1]), b.x, b.y);
                w[i] = (epi == EPI_RELU) ? pack_bf16x2_relu(s0, s1) : pack_bf16x2(s0, s1);
              }
            } else {
              float h[32];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 b = cb2[i];
                h[2 * i] = fmaxf(__uint_as_float(vc[2 * i]) + b.x, 0.f);
                h[2 * i + 1] = fmaxf(__uint_as_float(vc[2 * i + 1]) + b.y, 0.f);
                w[i] = pack_bf16x2(h[2 * i], h[2 * i + 1]);
              }
              if (epi == EPI_RELU_SIGMA) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sigma = fmaf(h[i], c_small[kSmallSigmaW + c0 + i], sigma);
              } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    rgb_acc[ch] = fmaf(h[i], c_small[kSmallRgbW + ch * 128 + c0 + i], rgb_acc[ch]);
                }
              }
            }
            // bf16 -> the next A tile (chunk = 64 columns = 128 B per row)
            const uint32_t chunk = sbase + kSmemAct + (c0 >> 6) * kChunkBytes;
            const int u0 = (c0 & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(chunk + sw128_off(row, u0 + j), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
        }
        if (epi == EPI_RELU_SIGMA) sigma += c_small[kSmallSigmaB];
        tc_fence_before();
        fence_proxy_async_smem();
        const bool last = (g == n_gemm - 1);
        if (!last) mbar_arrive(bar_a_ready);
        if (et == 0) FS_TRACE(titer, g, 5);
        if (stash_tile) {
          named_bar_sync(1, 128);
          if (et == 0) {
            bulk_s2g(stash_tile + L.stash_off, sbase + kSmemAct, (ncols >> 6) * kChunkBytes);
            if (epi == EPI_CONN)
              bulk_s2g(stash_tile + prog.stash_aux_dir_off, sbase + kSmemAux, kChunkBytes);
            bulk_commit();
          }
        }
        if (g == dir_after && !args.density_only) {
          // view-direction encoding for the branch layer, overlapped with the next layer's
          // MMAs: aux is free (its last reader, this layer's MMA, has completed) and the
          // fence.proxy.async of the following epilogues publishes it before the branch MMA
          encode_row(dir, prog.n_freqs_dir, prog.freq_dir, prog.pow2_freqs != 0, args.mask_dir, sbase + kSmemAux, row);
        }
        if (last && valid) {
          if (args.density_only) {
            args.out[args.density_only == 2 ? 4 * p + 3 : p] = sigma;
          } else {
            float r3[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              float z = rgb_acc[ch] + c_small[kSmallRgbB + ch];
              r3[ch] = 1.0f / (1.0f + expf(-z));
            }
            reinterpret_cast<float4*>(args.out)[p] = make_float4(r3[0], r3[1], r3[2], sigma);
          }
        }
      }
    }
    if (et == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace fs

using namespace fs;

extern "C" int fsnerf_mlp_forward(const fsnerf_net_cfg* cfg, const float* params,
                                  const void* packed, int64_t n_samples, int samples_per_ray,
                                  const float* rays_o, const float* rays_d, const float* t_starts,
                                  const float* t_ends, const float* x, const float* dirs,
                                  const float* mask_pos, const float* mask_dir, int density_only,
                                  float* out, void* stash, void* stream) {
  static MlpProgram P;
  int rc = build_program(cfg, &P);
  if (rc != FSNERF_OK) return rc;
  FS_REQUIRE(n_samples >= 0, "mlp_forward: negative n_samples");
  FS_REQUIRE(density_only >= 0 && density_only <= 2, "mlp_forward: density_only must be 0, 1 or 2");
  if (n_samples == 0) return FSNERF_OK;
  FS_REQUIRE(params && packed && out, "mlp_forward: null pointer");
  if (x) {
    FS_REQUIRE(density_only || dirs, "mlp_forward: dirs required unless density_only");
  } else {
    FS_REQUIRE(rays_o && rays_d && t_starts && t_ends && samples_per_ray > 0,
               "mlp_forward: rays/t_starts/t_ends required when x is NULL");
  }
  FS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0 &&
                 (reinterpret_cast<uintptr_t>(params) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(stash) & 127) == 0,
             "mlp_forward: params/out must be 16B aligned, packed/stash 128B aligned");
  FS_REQUIRE(!(stash && density_only), "mlp_forward: stash (training) needs the full network");
  if (n_samples == 0) return FSNERF_OK;
  static int variant = -1;
  if (variant < 0) {
    // FSNERF_FWD_VARIANT: unset / 2 = tensor-memory-resident kernel (mlp_fwd2.cu);
    // 0 / 1 = first-generation kernels kept for A/B measurements
    const char* e = getenv("FSNERF_FWD_VARIANT");
    variant = e ? atoi(e) : 2;
    cudaError_t e1 = cudaFuncSetAttribute(mlp_fwd_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem_total<2>());
    cudaError_t e2 = cudaFuncSetAttribute(mlp_fwd_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem_total<8>());
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      fsnerf_set_error("mlp_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      variant = -1;
      return FSNERF_ERR_CUDA;
    }
  }
  if (variant == 2)
    return mlp_forward_v2(P, packed, n_samples, samples_per_ray, rays_o, rays_d, t_starts, t_ends, x, dirs,
                          mask_pos, mask_dir, density_only, out, stash, stream);
  FwdArgs a;
  a.trace = reinterpret_cast<long long*>(fsnerf_debug_trace_ptr());
  a.params = params; a.packed = reinterpret_cast<const uint8_t*>(packed);
  a.n_samples = n_samples; a.samples_per_ray = samples_per_ray;
  a.rays_o = rays_o; a.rays_d = rays_d; a.t_starts = t_starts; a.t_ends = t_ends;
  a.x = x; a.dirs = dirs; a.mask_pos = mask_pos; a.mask_dir = mask_dir;
  a.density_only = density_only; a.out = out; a.stash = reinterpret_cast<uint8_t*>(stash);
  int64_t n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int ctas_per_sm = (variant == 1) ? 1 : 2;
  int grid = (int)(n_tiles < ctas_per_sm * kNumSMs ? n_tiles : ctas_per_sm * kNumSMs);
  {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_small, a.packed + P.small_off, kSmallFloats * sizeof(float), 0,
                                            cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) {
      fsnerf_set_error("mlp_forward: constant upload: %s", cudaGetErrorString(e));
      return FSNERF_ERR_CUDA;
    }
  }
  FsProfScope prof_(stash ? "mlp_fwd_train" : "mlp_fwd", stream);
  if (variant == 1)
    mlp_fwd_kernel<8, 1><<<grid, kFwdThreads, smem_total<8>(), (cudaStream_t)stream>>>(P, a);
  else
    mlp_fwd_kernel<2, 2><<<grid, kFwdThreads, smem_total<2>(), (cudaStream_t)stream>>>(P, a);
  return fsnerf_check_launch("mlp_forward");
}
