// Kernel (3) backward, second generation dgrad: the fused d(pre-activation) chain of the
// NeRF MLP with the gradients resident in TENSOR MEMORY (the loss.backward() edge of
// src/run-nerf.py:282 through src/core/models.py:111-143).  Same skeleton as mlp_fwd2.cu:
//   warps 0..7   epilogue: TMEM -> regs (loads pipelined one chunk ahead) -> (+ sigma-head
//                term) -> 1-bit ReLU mask written by the forward (32 B/sample/layer instead of
//                re-reading the 512 B activations) -> bf16x2 written back IN PLACE as the next step's A operand; every
//                finished chunk is handed to the MMA warps, staged through smem per 32-row
//                slab and bulk-stored to the dstash image that wgrad streams back.
//   warps 8..11  reducers: column sums of each staged slab = bias gradients (smem atomics)
//   warps 12,13  MMA issuers, alternating chunks (mlp_issue.cuh): D[128 x 256] = dpre . W
//                with W^T operand stages; the two TMEM regions alternate roles per step
//   warps 14,15  weight producers
// The seed step (d(out) -> rgb head^T -> branch-layer dpre) runs on CUDA cores in the
// epilogue warps.
#include <stdlib.h>
#include "common.cuh"
#include "mlp_common.cuh"
#include "mlp_issue.cuh"

namespace fs {
namespace {

constexpr int kEpiWarpsB = 8;
constexpr int kRedWarpsB = 4;
constexpr int kWarpRed0 = kEpiWarpsB;                 // 8
constexpr int kWarpMmaB = kEpiWarpsB + kRedWarpsB;    // 12, 13
constexpr int kWarpProdB = kWarpMmaB + kMmaWarps;     // 14, 15
constexpr int kThreadsB = (kWarpProdB + kProdWarps) * 32;  // 512
constexpr int kStagesB = 4;
constexpr int kSlabBytesB = 32 * 128;
constexpr int kStageBufsB = 3;
constexpr int kMaxLayersB = 12;

struct SmemB {
  static constexpr int ring = 0;
  static constexpr int staging = ring + kStagesB * kStageBytes;
  static constexpr int bias = staging + 4 * kStageBufsB * kSlabBytesB;   // fp32 [kMaxLayersB][256]
  static constexpr int heads = bias + kMaxLayersB * 256 * 4;             // sigma_w[256], rgb_w[3][128]
  static constexpr int bars = heads + 640 * 4;
  static constexpr int total = bars + 512;
};
struct BarsB {
  static constexpr int w_full = SmemB::bars;
  static constexpr int w_empty = w_full + 8 * kStagesB;
  static constexpr int a_ready = w_empty + 8 * kStagesB;  // [4]
  static constexpr int acc_full = a_ready + 8 * 4;        // [2]
  static constexpr int token = acc_full + 8 * 2;          // [2]
  static constexpr int slab_full = token + 16;            // [4 quarters][kStageBufsB]
  static constexpr int slab_free = slab_full + 8 * 4 * kStageBufsB;
  static constexpr int tmem_slot = slab_free + 8 * 4 * kStageBufsB;
};


struct StepB {
  int target;      // layer whose d(pre-activation) this step produces
  int mask_off;    // stash offset of the 1-bit ReLU mask of the target's forward output, -1: none
  int add_sigma;   // add d(sigma) * w_sigma (target is the last hidden layer)
  int dstash_off;  // where the target's dpre image goes in the backward record
};
struct PlanB {
  int n_steps;
  int g_branch, branch_mask_off, branch_dstash_off;
  StepB step[kMaxGemm];
};
struct ArgsB {
  long long* trace;
  const uint8_t* packed;
  int64_t n_samples;
  const uint8_t* stash;
  const float* out;
  const float* d_out;
  float* grads;
  uint8_t* dstash;
  int debug;  // tuning experiments only (FSNERF_DEBUG_FLAGS); results are wrong when non-zero
};

__global__ void __launch_bounds__(kThreadsB, 1)
mlp_dgrad2_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ PlanB plan,
                  const __grid_constant__ ArgsB args, const __grid_constant__ IssueTable tab) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_w_full = sbase + BarsB::w_full;
  const uint32_t bar_w_empty = sbase + BarsB::w_empty;
  const uint32_t bar_a_ready = sbase + BarsB::a_ready;
  const uint32_t bar_acc_full = sbase + BarsB::acc_full;
  const uint32_t bar_token = sbase + BarsB::token;
  const uint32_t bar_slab_full = sbase + BarsB::slab_full;
  const uint32_t bar_slab_free = sbase + BarsB::slab_free;
  const uint32_t tmem_slot = sbase + BarsB::tmem_slot;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + BarsB::tmem_slot);
  float* bias_acc = reinterpret_cast<float*>(smem + SmemB::bias);
  const float* heads = reinterpret_cast<const float*>(smem + SmemB::heads);
  const int64_t n_tiles = (args.n_samples + kTileM - 1) / kTileM;

  if ((sbase & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(bar_w_full + 8 * s, 1);
      mbar_init(bar_w_empty + 8 * s, 1);
    }
    for (int c = 0; c < 4; ++c) mbar_init(bar_a_ready + 8 * c, kEpiWarpsB);
    mbar_init(bar_acc_full, kMmaWarps);
    mbar_init(bar_acc_full + 8, kMmaWarps);
    mbar_init(bar_token, 1);
    mbar_init(bar_token + 8, 1);
    for (int i = 0; i < 4 * kStageBufsB; ++i) {
      mbar_init(bar_slab_full + 8 * i, 2);  // the two epilogue warps of the quarter
      mbar_init(bar_slab_free + 8 * i, 1);  // the quarter's store warp
    }
    fence_barrier_init();
  }
  if (warp == kWarpMmaB) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kMaxLayersB * 256; i += kThreadsB) bias_acc[i] = 0.f;
  {  // head weights straight from the packed image's fp32 small-params block (no constant upload)
    const float* __restrict__ small = reinterpret_cast<const float*>(args.packed + prog.small_off);
    for (int i = threadIdx.x; i < 640; i += kThreadsB)
      reinterpret_cast<float*>(smem + SmemB::heads)[i] = __ldg(small + kSmallSigmaW + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= kWarpProdB) {
    IssueBars IB{bar_w_full, bar_w_empty, bar_token, sbase + SmemB::ring};
    producer_loop<kStagesB>(tab, IB, args.packed, n_tiles, warp - kWarpProdB, lane);
  } else if (warp >= kWarpMmaB) {
    if (tmem_base != 0) __trap();
    IssueBars IB{bar_w_full, bar_w_empty, bar_token, sbase + SmemB::ring};
    issuer_loop<kStagesB>(tab, IB, sbase, n_tiles, (uint32_t)(warp - kWarpMmaB), lane, args.trace);
  } else if (warp >= kWarpRed0) {
    // ------------------------------------------------ store warps: dstash + bias gradients
    // one warp per lane quarter.  Per staged slab (32 rows x 64 features of one chunk): bulk
    // store it to the dstash image, sum its 32 rows (lane l owns features 2l, 2l+1:
    // conflict-free 4 B reads) into the bias-gradient accumulators, and free the buffer once
    // the bulk store has read it.  The epilogue warps never block on this.
    const int quarter = warp - kWarpRed0;
    const uint32_t stage_base = sbase + SmemB::staging + quarter * (kStageBufsB * kSlabBytesB);
    uint32_t n_staged = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // debug & 32: always the CTA's first record (stays L2 resident): same instructions, no HBM
      // writes.  Measured: no change (1.76 ms), while skipping the copy (debug & 4) gives 1.53 ms
      // and a second store warp per quarter changes nothing either — the copy costs through its
      // shared-memory / LSU traffic inside the SM, not through HBM or store-warp throughput.
      uint8_t* dstash_tile = args.dstash + (size_t)((args.debug & 32) ? (int64_t)blockIdx.x : tile) * prog.dstash_tile_bytes;
      for (int s = -1; s < plan.n_steps; ++s) {
        const int layer = (s < 0) ? plan.g_branch : plan.step[s].target;
        const int nchunk = (s < 0) ? 2 : 4;
        const int doff = (s < 0) ? plan.branch_dstash_off : plan.step[s].dstash_off;
        for (int c = 0; c < nchunk; ++c, ++n_staged) {
          const uint32_t b = n_staged % kStageBufsB;
          const uint32_t buf = stage_base + b * kSlabBytesB;
          mbar_wait_relaxed(bar_slab_full + 8 * (quarter * kStageBufsB + b), (n_staged / kStageBufsB) & 1);
          if (!(args.debug & 4)) {
            // coalesced copy with plain loads/stores (512 B per warp instruction): the epilogue
            // then needs no generic->async proxy fence (a MEMBAR.ALL.CTA per chunk) to hand over
            uint4* dst = reinterpret_cast<uint4*>(dstash_tile + doff + c * kChunkBytes + quarter * kSlabBytesB);
            uint4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(t[k].x), "=r"(t[k].y), "=r"(t[k].z), "=r"(t[k].w)
                           : "r"(buf + (k * 32 + lane) * 16));
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k * 32 + lane] = t[k];
          }
          float s0 = 0.f, s1 = 0.f;
          if (!(args.debug & 2))
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];"
                         : "=r"(w)
                         : "r"(buf + r * 128 + ((((uint32_t)lane >> 2) ^ (uint32_t)(r & 7)) << 4) + ((lane & 3) << 2)));
            s0 += bf16_lo(w);
            s1 += bf16_hi(w);
          }
          atomicAdd(bias_acc + layer * 256 + 64 * c + 2 * lane, s0);
          atomicAdd(bias_acc + layer * 256 + 64 * c + 2 * lane + 1, s1);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_slab_free + 8 * (quarter * kStageBufsB + b));
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stage_base = sbase + SmemB::staging + quarter * (kStageBufsB * kSlabBytesB);
    uint32_t acc_phase[2] = {0, 0};
    uint32_t n_staged = 0;
    uint32_t titer = 0;
    // stage the 32 bf16 of this thread (16 words) into the quarter's slab and hand it to the
    // store warp (SW128 image: unit u of row r at r*128 + ((u ^ (r&7)) << 4))
    auto stage_wait = [&]() {  // the buffer of the upcoming slab has been drained (3 slabs ago)
      const uint32_t b = n_staged % kStageBufsB;
      mbar_wait(bar_slab_free + 8 * (quarter * kStageBufsB + b), ((n_staged / kStageBufsB) & 1) ^ 1);
    };
    auto stage_out = [&](const uint32_t (&w)[16]) {
      const uint32_t b = n_staged % kStageBufsB;
      const uint32_t buf = stage_base + b * kSlabBytesB;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(buf + lane * 128 + (((uint32_t)(4 * half + j) ^ (uint32_t)(lane & 7)) << 4), w[4 * j],
                     w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_slab_full + 8 * (quarter * kStageBufsB + b));  // release: stores visible
      ++n_staged;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
      const int64_t p = tile * kTileM + row;
      const bool valid = p < args.n_samples;
      const uint8_t* stash_tile = args.stash + (size_t)tile * prog.stash_tile_bytes;
      // ---- seed: d(out) -> rgb head^T -> d(pre-activation) of the branch layer (128 wide)
      float dz[3] = {0.f, 0.f, 0.f}, dsig = 0.f;
      if (valid) {
        const float4 o4 = __ldg(reinterpret_cast<const float4*>(args.out) + p);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(args.d_out) + p);
        dz[0] = g4.x * o4.x * (1.0f - o4.x);  // sigmoid'
        dz[1] = g4.y * o4.y * (1.0f - o4.y);
        dz[2] = g4.z * o4.z * (1.0f - o4.z);
        dsig = g4.w;
      }
      {
        // ReLU masks are the 1-bit words the forward wrote (mlp_issue.cuh: relu_bits_*)
        const uint8_t* bm = stash_tile + plan.branch_mask_off;
        uint32_t bw = __ldg(reinterpret_cast<const uint32_t*>(bm + relu_bits_word_off(0, half, row)));
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int c0 = 64 * c + 32 * half;
          const uint32_t bw_next =
              (c == 0) ? __ldg(reinterpret_cast<const uint32_t*>(bm + relu_bits_word_off(1, half, row))) : 0u;
          uint32_t w[16];
          const float* wr = heads + 256 + c0;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int i = 2 * q;
            const float v0 = dz[0] * wr[i] + dz[1] * wr[128 + i] + dz[2] * wr[256 + i];
            const float v1 = dz[0] * wr[i + 1] + dz[1] * wr[128 + i + 1] + dz[2] * wr[256 + i + 1];
            w[q] = pack_bf16x2(v0, v1) & relu_bits_mask2(bw, q);
          }
          // step 0 reads its A operand from region 1
          tmem_st16(tmem_lane + 256u + c0, w);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_a_ready + 8 * c);
          stage_wait();
          stage_out(w);
          bw = bw_next;
        }
      }
      // ---- chain
      for (int s = 0; s < plan.n_steps; ++s) {
        const StepB S = plan.step[s];
        const int r = s & 1;
        const uint32_t region = tmem_lane + (uint32_t)r * 256u + 32u * half;
        const bool last = (s == plan.n_steps - 1);
        const uint8_t* mimg = (S.mask_off >= 0 && !(args.debug & 1)) ? stash_tile + S.mask_off : nullptr;
        const bool add_sigma = S.add_sigma != 0;
        // chunk 0's mask does not depend on the MMAs: fetch it before waiting on the accumulator
        uint32_t mw = mimg ? __ldg(reinterpret_cast<const uint32_t*>(mimg + relu_bits_word_off(0, half, row))) : 0xFFFFFFFFu;
        if (threadIdx.x == 0 && args.trace && blockIdx.x == 0 && titer < 4) args.trace[(titer * 16 + s) * 8 + 3] = clock64();
        mbar_wait(bar_acc_full + 8 * r, acc_phase[r]);
        acc_phase[r] ^= 1;
        tc_fence_after();
        if (threadIdx.x == 0 && args.trace && blockIdx.x == 0 && titer < 4) args.trace[(titer * 16 + s) * 8 + 4] = clock64();
        // The accumulator loads are software-pipelined one chunk ahead through two register
        // buffers: chunk c+1 is in flight from tensor memory while chunk c is converted, masked,
        // written back as the next step's A operand and staged for the dstash.
        auto chunk = [&](const int c, uint32_t (&v)[32], uint32_t (&vn)[32]) {
          tmem_ld_wait();  // v = chunk c
          if (c < 3) tmem_ld32(region + 64u * (c + 1), vn);
          const uint32_t mwn = (mimg && c < 3)
                                   ? __ldg(reinterpret_cast<const uint32_t*>(mimg + relu_bits_word_off(c + 1, half, row)))
                                   : 0xFFFFFFFFu;
          uint32_t w[16];
          if (add_sigma) {
            const float* ws = heads + 64 * c + 32 * half;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws + 4 * i);
              v[4 * i] = __float_as_uint(fmaf(dsig, w4.x, __uint_as_float(v[4 * i])));
              v[4 * i + 1] = __float_as_uint(fmaf(dsig, w4.y, __uint_as_float(v[4 * i + 1])));
              v[4 * i + 2] = __float_as_uint(fmaf(dsig, w4.z, __uint_as_float(v[4 * i + 2])));
              v[4 * i + 3] = __float_as_uint(fmaf(dsig, w4.w, __uint_as_float(v[4 * i + 3])));
            }
          }
#pragma unroll
          for (int q = 0; q < 16; ++q)
            w[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])) & relu_bits_mask2(mw, q);
          if (!last) {
            tmem_st16(region + 64u * c, w);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_ready + 8 * c);
          }
          stage_wait();
          stage_out(w);
          mw = mwn;
        };
        uint32_t va[32], vb[32];
        tmem_ld32(region, va);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          chunk(c, va, vb);
          chunk(c + 1, vb, va);
        }
        if (threadIdx.x == 0 && args.trace && blockIdx.x == 0 && titer < 4) args.trace[(titer * 16 + s) * 8 + 5] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMmaB) tmem_dealloc(tmem_base, 512);
  // bias gradients of this CTA -> global
  for (int g = 0; g < prog.n_gemm; ++g) {
    const int ncols = prog.layer[g].n_halves * 128;
    if ((int)threadIdx.x < ncols)
      atomicAdd(args.grads + prog.layer[g].bias_off + threadIdx.x, bias_acc[g * 256 + threadIdx.x]);
  }
}

}  // namespace

int mlp_dgrad_v2(const MlpProgram& P, const void* packed, int64_t n_samples, const void* stash, const float* out,
                 const float* d_out, float* grads, void* workspace, void* stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_dgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemB::total);
    if (e != cudaSuccess) {
      fsnerf_set_error("mlp_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return FSNERF_ERR_CUDA;
    }
    configured = true;
  }
  FS_REQUIRE(P.n_gemm <= kMaxLayersB, "mlp_backward: at most %d GEMM layers are supported", kMaxLayersB);
  static PlanB PL;
  static IssueTable T;
  PL.n_steps = P.n_gemm - 1;
  PL.g_branch = P.n_gemm - 1;
  PL.branch_mask_off = P.layer[P.n_gemm - 1].mask_off;
  PL.branch_dstash_off = P.layer[P.n_gemm - 1].dstash_off;
  int j = 0;
  int n_cons[4] = {0, 0, 0, 0};
  for (int s = 0; s < PL.n_steps; ++s)
    for (int c = 0; c < P.layer[P.n_gemm - 1 - s].bwd_n_chunks; ++c) ++n_cons[c];
  int cons[4] = {0, 0, 0, 0};
  for (int s = 0; s < PL.n_steps; ++s) {
    const int src = P.n_gemm - 1 - s, tgt = src - 1;
    StepB& S = PL.step[s];
    S.target = tgt;
    S.mask_off = P.layer[tgt].mask_off;
    S.add_sigma = (P.layer[tgt].epi == EPI_RELU_SIGMA) ? 1 : 0;
    S.dstash_off = P.layer[tgt].dstash_off;
    const int nch = P.layer[src].bwd_n_chunks;
    FS_REQUIRE(P.layer[src].bwd_n_halves == 2 && nch <= 4 && j + nch <= kMaxChunks2,
               "mlp_backward: unsupported layer shape for the tensor-memory dgrad");
    for (int c = 0; c < nch; ++c, ++j) {
      IssueRec& R = T.rec[j];
      // a_ready[c] completes once per producer of chunk c (seed or a step's epilogue) and is
      // consumed once per step with more than c chunks: index = tile_iter * n_cons[c] + cons[c]
      R.flags = ((uint32_t)s << 24) | ((s & 1) ? kRecDcol : 0u) | (c == 0 ? kRecFirst : 0u) |
                (c == nch - 1 ? kRecLast : 0u) | kRecTmem | ((uint32_t)(cons[c] & 1) << kRecParShift) |
                ((n_cons[c] & 1) ? kRecParTile : 0u);
      R.idesc = umma_idesc_bf16(128, 256, 0, 0);
      R.a0 = (uint32_t)((s + 1) & 1) * 256u + 64u * c;
      R.abar = BarsB::a_ready + 8 * c;
      R.xbar = 0;
      R.accbar = BarsB::acc_full + 8 * (s & 1);
      R.n_acc = issue_n_acc(c, nch);
      R.w_block = (uint32_t)(P.layer[src].bwd_first_block + c * 2);
      R.w_bytes = kStageBytes;
      ++cons[c];
    }
  }
  T.n = j;
  T.last_acc_off = BarsB::acc_full + 8 * ((PL.n_steps - 1) & 1);
  T.last_acc_n = 0;
  for (int s = 0; s < PL.n_steps; ++s)
    if ((s & 1) == ((PL.n_steps - 1) & 1)) ++T.last_acc_n;
  ArgsB a;
  a.trace = reinterpret_cast<long long*>(fsnerf_debug_trace_ptr());
  a.packed = reinterpret_cast<const uint8_t*>(packed);
  a.n_samples = n_samples;
  a.stash = reinterpret_cast<const uint8_t*>(stash);
  a.out = out; a.d_out = d_out; a.grads = grads;
  a.dstash = reinterpret_cast<uint8_t*>(workspace);
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("FSNERF_DEBUG_FLAGS"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
  }
  const int64_t n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  FsProfScope prof_("mlp_dgrad", stream);
  mlp_dgrad2_kernel<<<grid, kThreadsB, SmemB::total, (cudaStream_t)stream>>>(P, PL, a, T);
  return fsnerf_check_launch("mlp_backward(dgrad)");
}

}  // namespace fs
