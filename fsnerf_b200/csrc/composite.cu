// Kernel (4): alpha compositing forward/backward — a per-ray segmented
// transmittance scan.  HBM-bound: one warp per ray, lane-strided so that every
// global access is a fully coalesced 128 B (t) or 512 B (raw float4) line;
// the prefix (forward) and suffix (backward) scans run in registers with
// warp shuffles and a running carry across 32-sample blocks.
//
// Semantics (flags == 0): nerfacc.volrend.rendering v0.5.3 as called at
// /root/reference/src/render/rendering.py:89-96;
// oracle/compositing.py:composite_dense restates it.
#include "common.cuh"
#include "../../include/fsnerf_b200.h"

namespace {

constexpr float kEpsF32 = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_incl_scan_add(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float up = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += up;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float up = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v *= up;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// suffix (reverse) inclusive scan: lane l gets sum_{k>=l} v_k
__device__ __forceinline__ float warp_suffix_scan_add(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float dn = __shfl_down_sync(0xffffffffu, v, d);
    if (lane + d < 32) v += dn;
  }
  return v;
}

// Occlusion regulariser of src/core/loss.py:26-60 as called at src/run-nerf.py:260-264, fused
// into the compositing backward: loss = mean_r sum_s w(t_mid) * sigma_raw, so
// d(loss)/d(sigma_raw) = scale * w(t_mid) with scale = 1/(rays in the GLOBAL batch).
// func: 0 off, 1 'linear' w = -a t + b, 2 'exp' w = a exp(-b t).  loss (may be NULL)
// accumulates the UNSCALED sum over this launch's rays.
struct OccReg { int func; float a, b, scale; float* loss; };
__device__ __forceinline__ float occ_weight(const OccReg& o, float t) {
  return o.func == 1 ? fmaf(-o.a, t, o.b) : o.a * expf(-o.b * t);
}

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// NB = number of 32-sample blocks held in registers at once
template <int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(int64_t n_rays, int S, const float4* __restrict__ raw,
                     const float* __restrict__ ts, const float* __restrict__ te,
                     const float* __restrict__ dscale, const float* __restrict__ bkgd, int flags,
                     float* __restrict__ rgb, float* __restrict__ opacity,
                     float* __restrict__ depth, float* __restrict__ weights,
                     float* __restrict__ alphas, float* __restrict__ trans) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  const int64_t base = r * S;
  const float ds = dscale ? dscale[r] : 1.0f;
  const bool relu = flags & FSNERF_COMP_SIGMA_RELU;
  const bool prod = flags & FSNERF_COMP_PRODUCT_TRANS;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_d = 0.f;
  float carry = prod ? 1.0f : 0.0f;  // running exclusive sum of sigma*delta (or product)
  for (int s0 = 0; s0 < S; s0 += 32 * NB) {
    float4 c[NB];
    float t0[NB], t1[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      int s = s0 + b * 32 + lane;
      if (s < S) {
        c[b] = ld_stream_f4(raw + base + s);
        t0[b] = ld_stream_f(ts + base + s);
        t1[b] = ld_stream_f(te + base + s);
      } else {
        c[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        t0[b] = t1[b] = 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      int s = s0 + b * 32 + lane;
      if (s0 + b * 32 >= S) break;  // warp-uniform
      float sig = relu ? fmaxf(c[b].w, 0.f) : c[b].w;
      float sd = sig * ((t1[b] - t0[b]) * ds);
      float alpha = 1.0f - expf(-sd);
      float T;
      if (!prod) {
        float incl = warp_incl_scan_add(sd, lane);
        T = expf(-(carry + (incl - sd)));
        carry += __shfl_sync(0xffffffffu, incl, 31);
      } else {
        float f = 1.0f - alpha + 1e-10f;
        float incl = warp_incl_scan_mul(f, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        T = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
      float w = T * alpha;
      if (s < S) {
        weights[base + s] = w;
        if (alphas) alphas[base + s] = alpha;
        if (trans) trans[base + s] = T;
        acc_r += w * c[b].x;
        acc_g += w * c[b].y;
        acc_b += w * c[b].z;
        acc_w += w;
        acc_d += w * ((t0[b] + t1[b]) * 0.5f);
      }
    }
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
  acc_w = warp_sum(acc_w); acc_d = warp_sum(acc_d);
  if (lane == 0) {
    if (!(flags & FSNERF_COMP_DEPTH_UNNORM)) acc_d = acc_d / fmaxf(acc_w, kEpsF32);
    if (bkgd) {
      float om = 1.0f - acc_w;
      acc_r += bkgd[0] * om; acc_g += bkgd[1] * om; acc_b += bkgd[2] * om;
    }
    rgb[r * 3 + 0] = acc_r; rgb[r * 3 + 1] = acc_g; rgb[r * 3 + 2] = acc_b;
    opacity[r] = acc_w;
    depth[r] = acc_d;
  }
}

// Backward.  With sd_i = sigma_i*delta_i, w_i = T_i*alpha_i:
//   g_i   = dL/dw_i = d_rgb.c_i + dA + dDn*m_i + d_weights_i
//   exp-sum form:  dL/dsd_i = g_i*T_{i+1} - sum_{j>i} g_j w_j
//   product form:  dL/dsd_i = (1-a_i) * (g_i*T_i - sum_{j>i} g_j w_j / (1-a_i+1e-10))
// The whole ray is held in registers (NB blocks); S > 32*NB is rejected host-side.
template <int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (NB <= 6) ? 3 : 1)
composite_bwd_kernel(int64_t n_rays, int S, const float4* __restrict__ raw,
                     const float* __restrict__ ts, const float* __restrict__ te,
                     const float* __restrict__ dscale, const float* __restrict__ bkgd, int flags,
                     const float* __restrict__ d_rgb, const float* __restrict__ d_opacity,
                     const float* __restrict__ d_depth, const float* __restrict__ d_weights,
                     float4* __restrict__ d_raw, float* __restrict__ d_bkgd, const OccReg occ) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  const int64_t base = r * S;
  const float ds = dscale ? dscale[r] : 1.0f;
  const bool relu = flags & FSNERF_COMP_SIGMA_RELU;
  const bool prod = flags & FSNERF_COMP_PRODUCT_TRANS;
  float occ_acc = 0.f;
  float4 c[NB];
  float delta[NB], tm[NB], T[NB], al[NB], dw[NB];
  float carry = prod ? 1.0f : 0.0f;
  float acc_w = 0.f, acc_d = 0.f;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    int s = b * 32 + lane;
    if (s < S) {
      c[b] = ld_stream_f4(raw + base + s);
      float t0 = ld_stream_f(ts + base + s), t1 = ld_stream_f(te + base + s);
      delta[b] = (t1 - t0) * ds;
      tm[b] = (t0 + t1) * 0.5f;
      dw[b] = d_weights ? ld_stream_f(d_weights + base + s) : 0.f;
    } else {
      c[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      delta[b] = tm[b] = dw[b] = 0.f;
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b * 32 >= S) break;
    float sig = relu ? fmaxf(c[b].w, 0.f) : c[b].w;
    float sd = sig * delta[b];
    al[b] = 1.0f - expf(-sd);
    if (!prod) {
      float incl = warp_incl_scan_add(sd, lane);
      T[b] = expf(-(carry + (incl - sd)));
      carry += __shfl_sync(0xffffffffu, incl, 31);
    } else {
      float f = 1.0f - al[b] + 1e-10f;
      float incl = warp_incl_scan_mul(f, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      T[b] = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    float w = T[b] * al[b];
    if (b * 32 + lane < S) {
      acc_w += w;
      acc_d += w * tm[b];
    }
  }
  acc_w = warp_sum(acc_w);
  acc_d = warp_sum(acc_d);
  const float gr = d_rgb[r * 3], gg = d_rgb[r * 3 + 1], gb = d_rgb[r * 3 + 2];
  const float gA = d_opacity ? d_opacity[r] : 0.f;
  const float gD = d_depth ? d_depth[r] : 0.f;
  float dA = gA, dDn = gD;
  if (!(flags & FSNERF_COMP_DEPTH_UNNORM)) {
    float den = fmaxf(acc_w, kEpsF32);
    dDn = gD / den;
    if (acc_w > kEpsF32) dA -= gD * acc_d / (den * den);
  }
  if (bkgd) {
    float dot = gr * bkgd[0] + gg * bkgd[1] + gb * bkgd[2];
    dA -= dot;
    if (d_bkgd && lane == 0) {
      float om = 1.0f - acc_w;
      atomicAdd(d_bkgd + 0, gr * om);
      atomicAdd(d_bkgd + 1, gg * om);
      atomicAdd(d_bkgd + 2, gb * om);
    }
  }
  float suffix_carry = 0.f;  // sum over later blocks of g_j w_j
#pragma unroll
  for (int b = NB - 1; b >= 0; --b) {
    if (b * 32 >= S) continue;
    int s = b * 32 + lane;
    float w = T[b] * al[b];
    float g = gr * c[b].x + gg * c[b].y + gb * c[b].z + dA + dDn * tm[b] + dw[b];
    float gw = (s < S) ? g * w : 0.f;
    float incl = warp_suffix_scan_add(gw, lane);
    float later = suffix_carry + (incl - gw);  // strictly after s
    suffix_carry += __shfl_sync(0xffffffffu, incl, 0);
    float dsd;
    if (!prod) {
      dsd = g * (T[b] - w) - later;  // T_{i+1} = T_i (1-alpha_i) = T_i - w_i
    } else {
      float om = 1.0f - al[b];
      dsd = om * (g * T[b] - later / (om + 1e-10f));
    }
    float dsig = dsd * delta[b];
    if (relu && c[b].w <= 0.f) dsig = 0.f;
    if (occ.func && s < S) {
      const float ow = occ_weight(occ, tm[b]);
      dsig = fmaf(occ.scale, ow, dsig);
      occ_acc = fmaf(ow, c[b].w, occ_acc);
    }
    if (s < S) __stcs(d_raw + base + s, make_float4(w * gr, w * gg, w * gb, dsig));  // write-once stream
  }
  if (occ.func && occ.loss) {
    occ_acc = warp_sum(occ_acc);
    if (lane == 0) atomicAdd(occ.loss, occ_acc);
  }
}

// Streaming backward for the exp-sum transmittance (the reference semantics and every flag
// combination without FSNERF_COMP_PRODUCT_TRANS).  Two passes over the ray instead of holding
// it in registers: pass A is the forward scan (per-block carry-ins, opacity and depth sums),
// pass B walks the blocks backwards, re-reads them (L1/L2 hits: a ray is 4.6 KB), recomputes
// T with the SAME prefix arithmetic as the forward kernel and runs the suffix scan.  Small
// register state -> 2x the resident warps of the register-resident kernel -> HBM stays busy.
template <int NB, bool kOcc>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
composite_bwd_stream_kernel(int64_t n_rays, int S, const float4* __restrict__ raw,
                            const float* __restrict__ ts, const float* __restrict__ te,
                            const float* __restrict__ dscale, const float* __restrict__ bkgd, int flags,
                            const float* __restrict__ d_rgb, const float* __restrict__ d_opacity,
                            const float* __restrict__ d_depth, const float* __restrict__ d_weights,
                            float4* __restrict__ d_raw, float* __restrict__ d_bkgd, const OccReg occ) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  const int64_t base = r * S;
  const float ds = dscale ? dscale[r] : 1.0f;
  const bool relu = flags & FSNERF_COMP_SIGMA_RELU;
  float cin[NB];
  float carry = 0.f, acc_w = 0.f, acc_d = 0.f;
  // ---- pass A: forward scan
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    cin[b] = carry;
    if (b * 32 >= S) continue;  // warp-uniform
    const int s = b * 32 + lane;
    float sd = 0.f, tmid = 0.f;
    if (s < S) {
      const float sigma = __ldg(reinterpret_cast<const float*>(raw + base + s) + 3);
      const float t0 = __ldg(ts + base + s), t1 = __ldg(te + base + s);
      sd = (relu ? fmaxf(sigma, 0.f) : sigma) * ((t1 - t0) * ds);
      tmid = (t0 + t1) * 0.5f;
    }
    const float incl = warp_incl_scan_add(sd, lane);
    const float T = expf(-(carry + (incl - sd)));
    carry += __shfl_sync(0xffffffffu, incl, 31);
    const float w = T * (1.0f - expf(-sd));
    if (s < S) {
      acc_w += w;
      acc_d += w * tmid;
    }
  }
  acc_w = warp_sum(acc_w);
  acc_d = warp_sum(acc_d);
  const float gr = d_rgb[r * 3], gg = d_rgb[r * 3 + 1], gb = d_rgb[r * 3 + 2];
  const float gA = d_opacity ? d_opacity[r] : 0.f;
  const float gD = d_depth ? d_depth[r] : 0.f;
  float dA = gA, dDn = gD;
  if (!(flags & FSNERF_COMP_DEPTH_UNNORM)) {
    const float den = fmaxf(acc_w, kEpsF32);
    dDn = gD / den;
    if (acc_w > kEpsF32) dA -= gD * acc_d / (den * den);
  }
  if (bkgd) {
    const float dot = gr * bkgd[0] + gg * bkgd[1] + gb * bkgd[2];
    dA -= dot;
    if (d_bkgd && lane == 0) {
      const float om = 1.0f - acc_w;
      atomicAdd(d_bkgd + 0, gr * om);
      atomicAdd(d_bkgd + 1, gg * om);
      atomicAdd(d_bkgd + 2, gb * om);
    }
  }
  // ---- pass B: backward over the blocks
  float suffix_carry = 0.f;  // sum over later blocks of g_j w_j
  float occ_acc = 0.f;
#pragma unroll
  for (int b = NB - 1; b >= 0; --b) {
    if (b * 32 >= S) continue;
    const int s = b * 32 + lane;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    float delta = 0.f, tmid = 0.f, dwv = 0.f;
    if (s < S) {
      c = __ldg(raw + base + s);
      const float t0 = __ldg(ts + base + s), t1 = __ldg(te + base + s);
      delta = (t1 - t0) * ds;
      tmid = (t0 + t1) * 0.5f;
      if (d_weights) dwv = ld_stream_f(d_weights + base + s);
    }
    const float sd = (relu ? fmaxf(c.w, 0.f) : c.w) * delta;
    const float alpha = 1.0f - expf(-sd);
    const float incl_sd = warp_incl_scan_add(sd, lane);
    const float T = expf(-(cin[b] + (incl_sd - sd)));
    const float w = T * alpha;
    const float g = gr * c.x + gg * c.y + gb * c.z + dA + dDn * tmid + dwv;
    const float gw = (s < S) ? g * w : 0.f;
    const float incl = warp_suffix_scan_add(gw, lane);
    const float later = suffix_carry + (incl - gw);  // strictly after s
    suffix_carry += __shfl_sync(0xffffffffu, incl, 0);
    const float dsd = g * (T - w) - later;  // T_{i+1} = T_i (1-alpha_i) = T_i - w_i
    float dsig = dsd * delta;
    if (relu && c.w <= 0.f) dsig = 0.f;
    if (kOcc && s < S) {  // compiled out of the plain kernel: its register count sets the occupancy
      const float ow = occ_weight(occ, tmid);
      dsig = fmaf(occ.scale, ow, dsig);
      occ_acc = fmaf(ow, c.w, occ_acc);
    }
    if (s < S) __stcs(d_raw + base + s, make_float4(w * gr, w * gg, w * gb, dsig));  // write-once stream
  }
  if (kOcc && occ.loss) {
    occ_acc = warp_sum(occ_acc);
    if (lane == 0) atomicAdd(occ.loss, occ_acc);
  }
}

}  // namespace

extern "C" int fsnerf_composite_forward(int64_t n_rays, int n_samples, const float* raw,
                                        const float* t_starts, const float* t_ends,
                                        const float* delta_scale, const float* bkgd, int flags,
                                        float* rgb, float* opacity, float* depth, float* weights,
                                        float* alphas, float* trans, void* stream) {
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(raw && t_starts && t_ends && rgb && opacity && depth && weights,
             "composite_forward: null pointer");
  FS_REQUIRE(n_samples >= 1 && n_rays >= 0, "composite_forward: bad sizes");
  FS_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "composite_forward: raw must be 16B aligned");
  if (n_rays == 0) return FSNERF_OK;
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  const float4* raw4 = reinterpret_cast<const float4*>(raw);
#define LAUNCH(NB)                                                                              \
  composite_fwd_kernel<NB><<<blocks, kWarpsPerBlock * 32, 0, st>>>(                             \
      n_rays, n_samples, raw4, t_starts, t_ends, delta_scale, bkgd, flags, rgb, opacity, depth, \
      weights, alphas, trans)
  FsProfScope prof_("composite_fwd", stream);
  if (n_samples <= 64) LAUNCH(2);
  else if (n_samples <= 128) LAUNCH(4);
  else if (n_samples <= 192) LAUNCH(6);
  else LAUNCH(8);
#undef LAUNCH
  return fsnerf_check_launch("composite_forward");
}

static int composite_backward_impl(int64_t n_rays, int n_samples, const float* raw,
                                   const float* t_starts, const float* t_ends,
                                   const float* delta_scale, const float* bkgd, int flags,
                                   const float* d_rgb, const float* d_opacity,
                                   const float* d_depth, const float* d_weights, float* d_raw,
                                   float* d_bkgd, const OccReg occ, void* stream);

extern "C" int fsnerf_composite_backward(int64_t n_rays, int n_samples, const float* raw,
                                         const float* t_starts, const float* t_ends,
                                         const float* delta_scale, const float* bkgd, int flags,
                                         const float* d_rgb, const float* d_opacity,
                                         const float* d_depth, const float* d_weights, float* d_raw,
                                         float* d_bkgd, void* stream) {
  const OccReg occ = {0, 0.f, 0.f, 0.f, nullptr};
  return composite_backward_impl(n_rays, n_samples, raw, t_starts, t_ends, delta_scale, bkgd, flags,
                                 d_rgb, d_opacity, d_depth, d_weights, d_raw, d_bkgd, occ, stream);
}

extern "C" int fsnerf_composite_backward_occ(int64_t n_rays, int n_samples, const float* raw,
                                             const float* t_starts, const float* t_ends,
                                             const float* delta_scale, const float* bkgd, int flags,
                                             const float* d_rgb, const float* d_opacity,
                                             const float* d_depth, const float* d_weights,
                                             float* d_raw, float* d_bkgd, int occ_func, float occ_a,
                                             float occ_b, float occ_scale, float* occ_loss_sum,
                                             void* stream) {
  FS_REQUIRE(occ_func >= 0 && occ_func <= 2, "composite_backward_occ: occ_func must be 0 (off), 1 (linear) or 2 (exp)");
  FS_REQUIRE(occ_a >= 0.f && occ_b >= 0.f, "composite_backward_occ: a and b should be non-negative");
  const OccReg occ = {occ_func, occ_a, occ_b, occ_scale, occ_loss_sum};
  return composite_backward_impl(n_rays, n_samples, raw, t_starts, t_ends, delta_scale, bkgd, flags,
                                 d_rgb, d_opacity, d_depth, d_weights, d_raw, d_bkgd, occ, stream);
}

static int composite_backward_impl(int64_t n_rays, int n_samples, const float* raw,
                                   const float* t_starts, const float* t_ends,
                                   const float* delta_scale, const float* bkgd, int flags,
                                   const float* d_rgb, const float* d_opacity,
                                   const float* d_depth, const float* d_weights, float* d_raw,
                                   float* d_bkgd, const OccReg occ, void* stream) {
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(raw && t_starts && t_ends && d_rgb && d_raw, "composite_backward: null pointer");
  FS_REQUIRE(n_samples >= 1 && n_samples <= 512, "composite_backward: n_samples must be in [1,512]");
  FS_REQUIRE(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(d_raw)) & 15) == 0,
             "composite_backward: raw/d_raw must be 16B aligned");
  if (n_rays == 0) return FSNERF_OK;
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  const float4* raw4 = reinterpret_cast<const float4*>(raw);
  float4* d_raw4 = reinterpret_cast<float4*>(d_raw);
#define LAUNCH(NB)                                                                               \
  composite_bwd_kernel<NB><<<blocks, kWarpsPerBlock * 32, 0, st>>>(                              \
      n_rays, n_samples, raw4, t_starts, t_ends, delta_scale, bkgd, flags, d_rgb, d_opacity,     \
      d_depth, d_weights, d_raw4, d_bkgd, occ)
  FsProfScope prof_("composite_bwd", stream);
#define LAUNCH_STREAM(NB)                                                                        \
  do {                                                                                           \
    if (occ.func)                                                                                \
      composite_bwd_stream_kernel<NB, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(             \
          n_rays, n_samples, raw4, t_starts, t_ends, delta_scale, bkgd, flags, d_rgb, d_opacity, \
          d_depth, d_weights, d_raw4, d_bkgd, occ);                                              \
    else                                                                                         \
      composite_bwd_stream_kernel<NB, false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(            \
          n_rays, n_samples, raw4, t_starts, t_ends, delta_scale, bkgd, flags, d_rgb, d_opacity, \
          d_depth, d_weights, d_raw4, d_bkgd, occ);                                              \
  } while (0)
  if (!(flags & FSNERF_COMP_PRODUCT_TRANS)) {
    if (n_samples <= 64) LAUNCH_STREAM(2);
    else if (n_samples <= 128) LAUNCH_STREAM(4);
    else if (n_samples <= 192) LAUNCH_STREAM(6);
    else if (n_samples <= 256) LAUNCH_STREAM(8);
    else LAUNCH_STREAM(16);
    return fsnerf_check_launch("composite_backward");
  }
#undef LAUNCH_STREAM
  if (n_samples <= 64) LAUNCH(2);
  else if (n_samples <= 128) LAUNCH(4);
  else if (n_samples <= 192) LAUNCH(6);
  else if (n_samples <= 256) LAUNCH(8);
  else LAUNCH(16);
#undef LAUNCH
  return fsnerf_check_launch("composite_backward");
}
