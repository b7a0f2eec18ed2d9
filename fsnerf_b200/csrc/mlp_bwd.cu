// Kernel (3) backward: the NeRF MLP gradient (the loss.backward() edge of
// src/run-nerf.py:282 through src/core/models.py:111-143).  Three launches:
//
//  dgrad  fused chain per 128-sample tile (same warp roles / ring / TMEM use as
//         the forward): d(out) -> rgb head^T -> branch -> connection -> hidden
//         n-1 .. 1.  Each step is D[128 x 256] = dpre[128 x K'] . W (W^T blocks
//         streamed by bulk copy); the epilogue adds the sigma-head term, applies
//         the ReLU mask read from the forward stash, writes the bf16 dpre tile
//         as the next A operand, bulk-stores its image for wgrad and column-sums
//         it for the bias gradient.
//  wgrad  split-K tensor-core GEMMs dW[N_out x K_in] += dpre^T . X over all
//         samples: both operands are the stashed [samples x features] SW128
//         images used MN-major; persistent CTAs own (layer, tile-range) entries,
//         accumulate in TMEM and flush once with fp32 atomics.
//  heads  SIMT column sums for the degenerate (N=1 / N=3) sigma and rgb heads.
#include <stdlib.h>
#include "common.cuh"
#include "mlp_common.cuh"

namespace fs {
namespace {

// =========================================================================== dgrad
constexpr int kStages = 2;
constexpr int kThreads = 192;
constexpr int kSmemAct = 0;                                    // 4 chunks x 16 KB
constexpr int kSmemBias = kSmemAct + 4 * kChunkBytes;          // float [kMaxGemm][256] = 16 KB
constexpr int kSmemRing = kSmemBias + kMaxGemm * 256 * 4;      // kStages x 16 KB
constexpr int kSmemBars = kSmemRing + kStages * kBlockBytes;
constexpr int kSmemTotal = kSmemBars + 128;
constexpr int kTmemCols = 256;

struct BwdStep {
  int first_block;  // W^T blocks of the source layer
  int n_chunks;     // K' / 64
  int target;       // layer whose d(pre-activation) this step produces
  int mask_off;     // stash offset of the target's forward output (ReLU mask), -1: none
  int add_sigma;    // add d(sigma) * w_sigma (target is the last hidden layer)
};
struct BwdPlan {
  int n_steps;
  int n_blocks;
  BwdStep step[kMaxGemm];
};
struct BwdArgs {
  const float* params;
  const uint8_t* packed;
  int64_t n_samples;
  const uint8_t* stash;
  const float* out;
  const float* d_out;
  float* grads;
  uint8_t* dstash;
};

// sigma / rgb head weights (uniform epilogue operands), see mlp_common.cuh kSmall*
__constant__ float c_small[kSmallFloats];

__device__ __forceinline__ uint4 ldg_u4(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// column sums of the bf16 tile in the act buffer -> bias accumulator (smem)
__device__ __forceinline__ void column_sums(uint32_t act, float* bias_row, int et, int ncols) {
  if (2 * et >= ncols) return;
  const uint32_t chunk = act + (et >> 5) * kChunkBytes;
  const uint32_t unit = (et & 31) >> 2, wiu = et & 3;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int r = 0; r < kTileM; ++r) {
    uint32_t w;
    asm volatile("ld.shared.b32 %0, [%1];"
                 : "=r"(w)
                 : "r"(chunk + r * 128 + (((unit ^ (r & 7)) << 4) | (wiu << 2))));
    s0 += bf16_lo(w);
    s1 += bf16_hi(w);
  }
  bias_row[2 * et] += s0;
  bias_row[2 * et + 1] += s1;
}

__global__ void __launch_bounds__(kThreads, 2)
mlp_dgrad_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ BwdPlan plan,
                 const __grid_constant__ BwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_w_full = sbase + kSmemBars;
  const uint32_t bar_w_empty = bar_w_full + 8 * kStages;
  const uint32_t bar_a_ready = bar_w_empty + 8 * kStages;
  const uint32_t bar_acc_full = bar_a_ready + 8;
  const uint32_t tmem_slot = bar_acc_full + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 8 * (2 * kStages + 2));
  float* bias_acc = reinterpret_cast<float*>(smem + kSmemBias);
  const int64_t n_tiles = (args.n_samples + kTileM - 1) / kTileM;

  if ((sbase & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_w_full + 8 * s, 1);
      mbar_init(bar_w_empty + 8 * s, 1);
    }
    mbar_init(bar_a_ready, 128);
    mbar_init(bar_acc_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kMaxGemm * 256; i += kThreads) bias_acc[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    uint32_t cnt = 0;
    const uint8_t* wt = args.packed + (size_t)prog.n_blocks_fwd * kBlockBytes;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int b = 0; b < plan.n_blocks; ++b, ++cnt) {
        const uint32_t stage = cnt % kStages, phase = (cnt / kStages) & 1;
        mbar_wait(bar_w_empty + 8 * stage, phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_w_full + 8 * stage, kBlockBytes);
          bulk_g2s(sbase + kSmemRing + stage * kBlockBytes, wt + (size_t)b * kBlockBytes,
                   kBlockBytes, bar_w_full + 8 * stage);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    uint32_t cnt = 0, a_phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int s = 0; s < plan.n_steps; ++s) {
        const BwdStep& S = plan.step[s];
        mbar_wait(bar_a_ready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
        for (int c = 0; c < S.n_chunks; ++c) {
          const uint32_t a_tile = sbase + kSmemAct + c * kChunkBytes;
          for (int nh = 0; nh < 2; ++nh, ++cnt) {
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
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int g_branch = prog.n_gemm - 1;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t p = tile * kTileM + row;
      const bool valid = p < args.n_samples;
      const uint8_t* stash_tile = args.stash + (size_t)tile * prog.stash_tile_bytes;
      uint8_t* dstash_tile = args.dstash + (size_t)tile * prog.dstash_tile_bytes;
      // ---- seed: d(out) -> rgb head^T -> d(pre-activation) of the branch layer
      float dz[3] = {0.f, 0.f, 0.f}, dsig = 0.f;
      if (valid) {
        const float4 o4 = __ldg(reinterpret_cast<const float4*>(args.out) + p);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(args.d_out) + p);
        dz[0] = g4.x * o4.x * (1.0f - o4.x);  // sigmoid'
        dz[1] = g4.y * o4.y * (1.0f - o4.y);
        dz[2] = g4.z * o4.z * (1.0f - o4.z);
        dsig = g4.w;
      }
      if (et == 0) bulk_wait_read0();
      named_bar_sync(1, 128);  // previous tile: image stores + column sums done
      {
        const GemmLayer& LB = prog.layer[g_branch];
        const uint8_t* hb = stash_tile + LB.stash_off;
        uint4 m[16];  // ReLU mask source: the branch layer's forward output row (128 bf16)
#pragma unroll
        for (int j = 0; j < 16; ++j)
          m[j] = ldg_u4(hb + (j >> 3) * kChunkBytes + sw128_off(row, j & 7));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t mw[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
          uint32_t pw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = 8 * j + 2 * q;
            float v0 = dz[0] * c_small[kSmallRgbW + col] + dz[1] * c_small[kSmallRgbW + 128 + col] +
                       dz[2] * c_small[kSmallRgbW + 256 + col];
            float v1 = dz[0] * c_small[kSmallRgbW + col + 1] + dz[1] * c_small[kSmallRgbW + 128 + col + 1] +
                       dz[2] * c_small[kSmallRgbW + 256 + col + 1];
            if ((mw[q] & 0x00007FFFu) == 0u) v0 = 0.f;  // relu'(h): h == 0 <=> masked
            if ((mw[q] & 0x7FFF0000u) == 0u) v1 = 0.f;
            pw[q] = pack_bf16x2(v0, v1);
          }
          st_shared_v4(sbase + kSmemAct + (j >> 3) * kChunkBytes + sw128_off(row, j & 7), pw[0], pw[1],
                       pw[2], pw[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_a_ready);
        named_bar_sync(1, 128);
        if (et == 0) {
          bulk_s2g(dstash_tile + LB.dstash_off, sbase + kSmemAct, 2 * kChunkBytes);
          bulk_commit();
        }
        column_sums(sbase + kSmemAct, bias_acc + g_branch * 256, et, 128);
      }
      // ---- chain
      for (int s = 0; s < plan.n_steps; ++s) {
        const BwdStep& S = plan.step[s];
        const GemmLayer& LT = prog.layer[S.target];
        const uint8_t* mimg = (S.mask_off >= 0) ? stash_tile + S.mask_off : nullptr;
        const bool add_sigma = S.add_sigma != 0;
        // masks do not depend on the MMA: fetch chunk 0's before waiting on the accumulator
        uint4 m[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          m[0][j] = mimg ? ldg_u4(mimg + sw128_off(row, j)) : make_uint4(~0u, ~0u, ~0u, ~0u);
        mbar_wait(bar_acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (et == 0) bulk_wait_read0();
        named_bar_sync(1, 128);  // act free: MMA, image store and column sums are done
        uint32_t v[2][32];
        tmem_ld32(tmem_row, v[0]);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c0 = ci * 32;
          tmem_ld_wait();
          if (ci < 7) {
            tmem_ld32(tmem_row + c0 + 32, v[(ci + 1) & 1]);
            const int c1 = c0 + 32;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              m[(ci + 1) & 1][j] = mimg ? ldg_u4(mimg + (c1 >> 6) * kChunkBytes + sw128_off(row, ((c1 & 63) >> 3) + j))
                                        : make_uint4(~0u, ~0u, ~0u, ~0u);
          }
          const uint32_t(&vc)[32] = v[ci & 1];
          const int u0 = (c0 & 63) >> 3;
          const uint32_t chunk = sbase + kSmemAct + (c0 >> 6) * kChunkBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 m4 = m[ci & 1][j];
            const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
            uint32_t pw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = 8 * j + 2 * q;
              float v0 = __uint_as_float(vc[i]), v1 = __uint_as_float(vc[i + 1]);
              if (add_sigma) {
                v0 = fmaf(dsig, c_small[kSmallSigmaW + c0 + i], v0);
                v1 = fmaf(dsig, c_small[kSmallSigmaW + c0 + i + 1], v1);
              }
              if ((mw[q] & 0x00007FFFu) == 0u) v0 = 0.f;
              if ((mw[q] & 0x7FFF0000u) == 0u) v1 = 0.f;
              pw[q] = pack_bf16x2(v0, v1);
            }
            st_shared_v4(chunk + sw128_off(row, u0 + j), pw[0], pw[1], pw[2], pw[3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        if (s + 1 < plan.n_steps) mbar_arrive(bar_a_ready);
        named_bar_sync(1, 128);
        if (et == 0) {
          bulk_s2g(dstash_tile + LT.dstash_off, sbase + kSmemAct, 4 * kChunkBytes);
          bulk_commit();
        }
        column_sums(sbase + kSmemAct, bias_acc + S.target * 256, et, 256);
      }
    }
    if (et == 0) bulk_wait0();
    named_bar_sync(1, 128);
    for (int g = 0; g < prog.n_gemm; ++g) {
      const int ncols = prog.layer[g].n_halves * 128;
      if (2 * et < ncols) {
        atomicAdd(args.grads + prog.layer[g].bias_off + 2 * et, bias_acc[g * 256 + 2 * et]);
        atomicAdd(args.grads + prog.layer[g].bias_off + 2 * et + 1, bias_acc[g * 256 + 2 * et + 1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =========================================================================== wgrad
constexpr int kWStages = 3;
constexpr int kSlabRows = 64;                     // samples per stage
constexpr int kSlabBytes = kSlabRows * 128;       // 4 KB per 64-feature chunk
constexpr int kWStageBytes = 8 * kSlabBytes;      // A: 4 chunks, B: 4 chunks
constexpr int kWSmemBars = kWStages * kWStageBytes;
constexpr int kWSmemTotal = kWSmemBars + 128;
constexpr int kWTmemCols = 512;

struct WgradJob {
  int a_off, a_chunks;  // dpre image (dstash record), N_out / 64
  int b_off, b_chunks;  // input image (stash record), K_in(part) / 64
  int w_off, ld, col0, ncols, nrows;
  int cta_begin, n_split;
};
struct WgradPlan {
  int n_jobs, n_ctas;
  WgradJob job[kMaxGemm + 4];
};
struct WgradArgs {
  const uint8_t* stash;
  const uint8_t* dstash;
  int stash_tile_bytes, dstash_tile_bytes;
  int64_t n_tiles;
  float* grads;
};

__global__ void __launch_bounds__(kThreads, 1)
mlp_wgrad_kernel(const __grid_constant__ WgradPlan plan, const __grid_constant__ WgradArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + kWSmemBars;
  const uint32_t bar_empty = bar_full + 8 * kWStages;
  const uint32_t bar_acc_full = bar_empty + 8 * kWStages;
  const uint32_t tmem_slot = bar_acc_full + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + kWSmemBars + 8 * (2 * kWStages + 1));
  int j = 0;
  while (j + 1 < plan.n_jobs && (int)blockIdx.x >= plan.job[j + 1].cta_begin) ++j;
  const WgradJob& J = plan.job[j];
  const int part = blockIdx.x - J.cta_begin;
  const int64_t t0 = args.n_tiles * part / J.n_split, t1 = args.n_tiles * (part + 1) / J.n_split;
  if ((sbase & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(bar_full + 8 * s, 8);   // eight issuing threads (two lanes of each of the four producer warps)
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kWTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int n_mh = J.a_chunks / 2;

  if (t1 > t0) {
    if (warp >= 2) {
      // four producer warps (the epilogue warps, idle during the main loop): bulk copies issued
      // by one thread do not overlap (tools/l2_bench.cu), so each stage's slab copies (8 KB
      // each: 64 rows of one chunk image) are spread over eight issuing threads, each arming
      // the stage barrier for its own bytes
      const int pw = (warp - 2) * 2 + lane;  // issuing thread index (lanes 0 and 1 issue)
      const int n_cp = J.a_chunks + J.b_chunks;
      uint32_t cnt = 0;
      for (int64_t tile = t0; tile < t1; ++tile) {
        const uint8_t* a_img = args.dstash + (size_t)tile * args.dstash_tile_bytes + J.a_off;
        const uint8_t* b_img = args.stash + (size_t)tile * args.stash_tile_bytes + J.b_off;
        for (int slab = 0; slab < kTileM / kSlabRows; ++slab, ++cnt) {
          const uint32_t stage = cnt % kWStages, phase = (cnt / kWStages) & 1;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (lane < 2) {
            const uint32_t sa = sbase + stage * kWStageBytes, sb = sa + 4 * kSlabBytes;
            int mine = 0;
            for (int c = pw; c < n_cp; c += 8) ++mine;
            mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)mine * kSlabBytes);
            for (int c = pw; c < n_cp; c += 8) {
              if (c < J.a_chunks)
                bulk_g2s(sa + c * kSlabBytes, a_img + c * kChunkBytes + slab * kSlabBytes, kSlabBytes,
                         bar_full + 8 * stage);
              else
                bulk_g2s(sb + (c - J.a_chunks) * kSlabBytes, b_img + (c - J.a_chunks) * kChunkBytes + slab * kSlabBytes,
                         kSlabBytes, bar_full + 8 * stage);
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // A = dpre^T (M = output features), B = X^T (N = input features); both MN-major:
      // 64-feature groups LBO = kSlabBytes apart, 8-sample K groups SBO = 1024 B apart.
      const uint32_t idesc = umma_idesc_bf16(128, J.b_chunks * 64, 1, 1);
      uint32_t cnt = 0;
      const int64_t n_stages = (t1 - t0) * (kTileM / kSlabRows);
      for (int64_t it = 0; it < n_stages; ++it, ++cnt) {
        const uint32_t stage = cnt % kWStages, phase = (cnt / kWStages) & 1;
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = sbase + stage * kWStageBytes, sb = sa + 4 * kSlabBytes;
          for (int mh = 0; mh < n_mh; ++mh) {
#pragma unroll
            for (int ks = 0; ks < kSlabRows / 16; ++ks) {
              umma_bf16_ss(tmem_base + mh * 256,
                           umma_desc_sw128(sa + mh * 2 * kSlabBytes + ks * 2048, kSlabBytes, 1024),
                           umma_desc_sw128(sb + ks * 2048, kSlabBytes, 1024), idesc,
                           (it > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(bar_acc_full);
      __syncwarp();
    }
    if (warp >= 2) {
      const int quarter = warp & 3;
      mbar_wait(bar_acc_full, 0);
      tc_fence_after();
      for (int mh = 0; mh < n_mh; ++mh) {
        const int r = mh * 128 + quarter * 32 + lane;
        float* __restrict__ grow = args.grads + J.w_off + (size_t)r * J.ld + J.col0;
        for (int c0 = 0; c0 < J.b_chunks * 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + mh * 256 + c0, v);
          tmem_ld_wait();
          if (r < J.nrows) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < J.ncols) atomicAdd(grow + c0 + i, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWTmemCols);
}

// =========================================================================== heads (SIMT)
struct HeadsArgs {
  const uint8_t* stash;
  int stash_tile_bytes;
  int h_off;   // stash offset of the last hidden layer's output image (sigma head input)
  int hb_off;  // stash offset of the branch output image (rgb head input)
  int64_t n_samples, n_tiles;
  const float* out;
  const float* d_out;
  float* g_sigma_w; float* g_sigma_b; float* g_rgb_w; float* g_rgb_b;
};

// 256 threads: thread t owns 16-byte unit (t & 31) [= 8 features of chunk (t&31)>>3]
// of the 256-wide h image and, if (t & 31) < 16, of the 128-wide hb image; the 8
// row-groups (t >> 5) split the 128 rows.  Every load is a coalesced 16 B / lane.
__global__ void __launch_bounds__(256, 3)
mlp_heads_wgrad_kernel(const __grid_constant__ HeadsArgs a) {
  __shared__ float4 dsm[kTileM];  // (dz0, dz1, dz2, dsigma) per row
  __shared__ float red[8][32][33];
  const int t = threadIdx.x, u = t & 31, rg = t >> 5;
  const int chunk = u >> 3, unit = u & 7;
  float acc_s[8], acc_r[3][8];
  float4 acc_b = make_float4(0.f, 0.f, 0.f, 0.f);  // bias sums: lane 0 of each row group
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc_s[e] = 0.f; acc_r[0][e] = acc_r[1][e] = acc_r[2][e] = 0.f; }
  // (dz, dsigma) of this thread's row of the NEXT tile, fetched one tile ahead so that its
  // latency overlaps the current tile's streaming loop
  auto load_d = [&](int64_t tile) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t p = tile * kTileM + t;
    if (t < kTileM && tile < a.n_tiles && p < a.n_samples) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out) + p);
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.d_out) + p);
      d = make_float4(g4.x * o4.x * (1.f - o4.x), g4.y * o4.y * (1.f - o4.y),
                      g4.z * o4.z * (1.f - o4.z), g4.w);
    }
    return d;
  };
  float4 d_next = load_d(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (t < kTileM) dsm[t] = d_next;
    __syncthreads();
    d_next = load_d(tile + gridDim.x);
    const uint8_t* rec = a.stash + (size_t)tile * a.stash_tile_bytes;
    const uint8_t* h = rec + a.h_off + chunk * kChunkBytes;
    // lanes with u >= 16 have no hb column: they re-read a valid unit (same cache lines as
    // lanes 0..15) and their rgb partial sums are never used, so the loop stays branch-free
    // and every load of a batch is in flight before the first use
    const uint8_t* hb = rec + a.hb_off + (chunk & 1) * kChunkBytes;
    constexpr int kBatch = 4;
#pragma unroll 1
    for (int r0 = 0; r0 < kTileM / 8; r0 += kBatch) {
      uint4 hv[kBatch], bv[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int r = rg * (kTileM / 8) + r0 + k;
        const uint32_t off = sw128_off(r, unit);
        hv[k] = ldg_u4(h + off);
        bv[k] = ldg_u4(hb + off);
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int r = rg * (kTileM / 8) + r0 + k;
        const float4 d = dsm[r];
        if (u == 0) { acc_b.x += d.x; acc_b.y += d.y; acc_b.z += d.z; acc_b.w += d.w; }
        const uint32_t hw[4] = {hv[k].x, hv[k].y, hv[k].z, hv[k].w};
        const uint32_t bw[4] = {bv[k].x, bv[k].y, bv[k].z, bv[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc_s[2 * q] = fmaf(d.w, bf16_lo(hw[q]), acc_s[2 * q]);
          acc_s[2 * q + 1] = fmaf(d.w, bf16_hi(hw[q]), acc_s[2 * q + 1]);
          const float x0 = bf16_lo(bw[q]), x1 = bf16_hi(bw[q]);
          acc_r[0][2 * q] = fmaf(d.x, x0, acc_r[0][2 * q]); acc_r[0][2 * q + 1] = fmaf(d.x, x1, acc_r[0][2 * q + 1]);
          acc_r[1][2 * q] = fmaf(d.y, x0, acc_r[1][2 * q]); acc_r[1][2 * q + 1] = fmaf(d.y, x1, acc_r[1][2 * q + 1]);
          acc_r[2][2 * q] = fmaf(d.z, x0, acc_r[2][2 * q]); acc_r[2][2 * q + 1] = fmaf(d.z, x1, acc_r[2][2 * q + 1]);
        }
      }
    }
  }
  // combine the 8 row-groups through shared memory, then one atomic per feature
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rg][u][e] = acc_s[e];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[rg][u][8 + 8 * c + e] = acc_r[c][e];
  __syncthreads();
  {
    // thread t -> feature t of h (sigma head): unit t>>3, element t&7
    float s = 0.f;
    for (int g = 0; g < 8; ++g) s += red[g][t >> 3][t & 7];
    atomicAdd(a.g_sigma_w + t, s);
    if (t < 128) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float r3 = 0.f;
        for (int g = 0; g < 8; ++g) r3 += red[g][t >> 3][8 + 8 * c + (t & 7)];
        atomicAdd(a.g_rgb_w + c * 128 + t, r3);
      }
    }
  }
  if (u == 0) {
    atomicAdd(a.g_rgb_b + 0, acc_b.x);
    atomicAdd(a.g_rgb_b + 1, acc_b.y);
    atomicAdd(a.g_rgb_b + 2, acc_b.z);
    atomicAdd(a.g_sigma_b, acc_b.w);
  }
}

}  // namespace
}  // namespace fs

using namespace fs;

extern "C" int64_t fsnerf_mlp_bwd_workspace_bytes(const fsnerf_net_cfg* cfg, int64_t n_samples) {
  MlpProgram P;
  if (build_program(cfg, &P) != FSNERF_OK) return -1;
  int64_t tiles = (n_samples + kTileM - 1) / kTileM;
  return tiles * (int64_t)P.dstash_tile_bytes;
}

extern "C" int fsnerf_mlp_backward(const fsnerf_net_cfg* cfg, const float* params,
                                   const void* packed, int64_t n_samples, const void* stash,
                                   const float* out, const float* d_out, int density_only,
                                   float* grads, void* workspace, void* stream) {
  static MlpProgram P;
  int rc = build_program(cfg, &P);
  if (rc != FSNERF_OK) return rc;
  if (density_only) {
    fsnerf_set_error("mlp_backward: density_only backward is not supported (the reference's sigma_fn "
                     "pass runs under no_grad, src/render/rendering.py:58-64)");
    return FSNERF_ERR_UNSUPPORTED;
  }
  FS_REQUIRE(n_samples >= 0, "mlp_backward: negative n_samples");
  if (n_samples == 0) return FSNERF_OK;
  FS_REQUIRE(params && packed && stash && out && d_out && grads && workspace,
             "mlp_backward: null pointer");
  FS_REQUIRE(((reinterpret_cast<uintptr_t>(stash) | reinterpret_cast<uintptr_t>(workspace) |
               reinterpret_cast<uintptr_t>(packed)) & 127) == 0 &&
                 ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_out) |
                   reinterpret_cast<uintptr_t>(params)) & 15) == 0,
             "mlp_backward: stash/workspace/packed must be 128B aligned, out/d_out/params 16B");
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    cudaError_t e2 = cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemTotal);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      fsnerf_set_error("mlp_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return FSNERF_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t n_tiles = (n_samples + kTileM - 1) / kTileM;
  // ---- dgrad plan
  static BwdPlan BP;
  BP.n_steps = P.n_gemm - 1;
  BP.n_blocks = P.n_blocks_bwd;
  for (int s = 0; s < BP.n_steps; ++s) {
    const int src = P.n_gemm - 1 - s, tgt = src - 1;
    BwdStep& S = BP.step[s];
    S.first_block = P.layer[src].bwd_first_block - P.n_blocks_fwd;
    S.n_chunks = P.layer[src].bwd_n_chunks;
    S.target = tgt;
    S.mask_off = (P.layer[tgt].epi == EPI_CONN) ? -1 : P.layer[tgt].stash_off;
    S.add_sigma = (P.layer[tgt].epi == EPI_RELU_SIGMA) ? 1 : 0;
  }
  BwdArgs ba;
  ba.params = params; ba.packed = reinterpret_cast<const uint8_t*>(packed); ba.n_samples = n_samples;
  ba.stash = reinterpret_cast<const uint8_t*>(stash); ba.out = out; ba.d_out = d_out;
  ba.grads = grads; ba.dstash = reinterpret_cast<uint8_t*>(workspace);
  int grid = (int)(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
  static int variant = -1;  // FSNERF_BWD_VARIANT: unset / 2 = tensor-memory dgrad (mlp_bwd2.cu), 1 = first generation
  if (variant < 0) {
    const char* e = getenv("FSNERF_BWD_VARIANT");
    variant = e ? atoi(e) : 2;
    // the tensor-memory dgrad reads the 1-bit ReLU masks that only the second-generation
    // forward writes into the stash
    const char* f = getenv("FSNERF_FWD_VARIANT");
    if (f && atoi(f) != 2) variant = 1;
  }
  if (variant == 2) {
    rc = mlp_dgrad_v2(P, packed, n_samples, stash, out, d_out, grads, workspace, stream);
  } else {
    {
      cudaError_t e = cudaMemcpyToSymbolAsync(c_small, ba.packed + P.small_off, kSmallFloats * sizeof(float), 0,
                                              cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) {
        fsnerf_set_error("mlp_backward: constant upload: %s", cudaGetErrorString(e));
        return FSNERF_ERR_CUDA;
      }
    }
    {
      FsProfScope prof_("mlp_dgrad", stream);
      mlp_dgrad_kernel<<<grid, kThreads, kSmemTotal, st>>>(P, BP, ba);
    }
    rc = fsnerf_check_launch("mlp_backward(dgrad)");
  }
  if (rc != FSNERF_OK) return rc;
  // ---- heads
  HeadsArgs ha;
  ha.stash = ba.stash; ha.stash_tile_bytes = P.stash_tile_bytes;
  ha.h_off = P.layer[P.n_hidden - 1].stash_off; ha.hb_off = P.layer[P.n_gemm - 1].stash_off;
  ha.n_samples = n_samples; ha.n_tiles = n_tiles; ha.out = out; ha.d_out = d_out;
  ha.g_sigma_w = grads + P.sigma_w_off; ha.g_sigma_b = grads + P.sigma_b_off;
  ha.g_rgb_w = grads + P.rgb_w_off; ha.g_rgb_b = grads + P.rgb_b_off;
  int hgrid = (int)(n_tiles < 6 * kNumSMs ? n_tiles : 6 * kNumSMs);
  {
    FsProfScope prof_("mlp_heads_wgrad", stream);
    mlp_heads_wgrad_kernel<<<hgrid, 256, 0, st>>>(ha);
  }
  rc = fsnerf_check_launch("mlp_backward(heads)");
  if (rc != FSNERF_OK) return rc;
  // ---- wgrad plan: (layer, input part) jobs, CTAs split proportionally to the bytes they stream
  static WgradPlan WP;
  WP.n_jobs = 0;
  double cost[kMaxGemm + 4], total = 0;
  for (int g = 0; g < P.n_gemm; ++g) {
    const GemmLayer& L = P.layer[g];
    const int a_chunks = L.n_halves * 2;
    for (int part = 0; part < 2; ++part) {
      if (part == 0 && L.n_act_chunks == 0) continue;
      if (part == 1 && !L.use_aux) continue;
      WgradJob& J = WP.job[WP.n_jobs];
      J.a_off = L.dstash_off; J.a_chunks = a_chunks;
      J.w_off = L.w_off; J.ld = L.ld; J.nrows = L.n_halves * 128;
      if (part == 0) {
        J.b_off = P.layer[g - 1].stash_off; J.b_chunks = L.n_act_chunks;
        J.col0 = 0; J.ncols = L.n_act_chunks * 64;
      } else {
        J.b_off = (L.epi == EPI_BRANCH) ? P.stash_aux_dir_off : P.stash_aux_pos_off;
        J.b_chunks = 1; J.col0 = L.n_act_chunks * 64; J.ncols = L.ld - J.col0;
      }
      cost[WP.n_jobs] = J.a_chunks + J.b_chunks;
      total += cost[WP.n_jobs];
      ++WP.n_jobs;
    }
  }
  // CTAs per job proportional to the bytes it streams; the kernel ends with its slowest job, so
  // the CTAs left over by rounding go, one at a time, to the job with the most bytes per CTA
  int n_cta[kMaxGemm + 4], used = 0;
  for (int jn = 0; jn < WP.n_jobs; ++jn) {
    int n = (int)(kNumSMs * cost[jn] / total);
    if (n < 1) n = 1;
    if ((int64_t)n > n_tiles) n = (int)n_tiles;
    n_cta[jn] = n;
    used += n;
  }
  while (used < kNumSMs) {
    int best = -1;
    for (int jn = 0; jn < WP.n_jobs; ++jn)
      if ((int64_t)n_cta[jn] < n_tiles && (best < 0 || cost[jn] / n_cta[jn] > cost[best] / n_cta[best])) best = jn;
    if (best < 0) break;
    ++n_cta[best];
    ++used;
  }
  int begin = 0;
  for (int jn = 0; jn < WP.n_jobs; ++jn) {
    WP.job[jn].cta_begin = begin;
    WP.job[jn].n_split = n_cta[jn];
    begin += n_cta[jn];
  }
  WP.n_ctas = begin;
  WgradArgs wa;
  wa.stash = ba.stash; wa.dstash = ba.dstash; wa.stash_tile_bytes = P.stash_tile_bytes;
  wa.dstash_tile_bytes = P.dstash_tile_bytes; wa.n_tiles = n_tiles; wa.grads = grads;
  FsProfScope prof_("mlp_wgrad", stream);
  mlp_wgrad_kernel<<<WP.n_ctas, kThreads, kWSmemTotal, st>>>(WP, wa);
  return fsnerf_check_launch("mlp_backward(wgrad)");
}
