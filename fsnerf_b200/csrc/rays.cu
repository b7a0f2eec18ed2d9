// Kernel family (1): ray generation (+NDC, +rgb gather), stratified sampling,
// hierarchical sample_pdf.  HBM/latency-bound integer + fp32 bookkeeping.
//
// Every fp32 product/sum whose rounding reaches an integer decision (pixel
// bookkeeping, searchsorted index, sort permutation) is written with the
// _rn intrinsics so nvcc cannot contract it into an FMA: oracle/rays.py and
// oracle/sampling.py restate the same op order in numpy and must agree bit
// for bit.
#include "common.cuh"
#include "../../include/fsnerf_b200.h"

namespace {

// reference: src/utils/utilities.py:54-82 (get_rays) and :102-120 (to_ndc)
__device__ __forceinline__ void ndc_warp(float& ox, float& oy, float& oz, float& dx, float& dy,
                                         float& dz, float near, float sx, float sy) {
  float t = __fdiv_rn(-__fadd_rn(near, oz), dz);
  ox = __fadd_rn(ox, __fmul_rn(t, dx));
  oy = __fadd_rn(oy, __fmul_rn(t, dy));
  oz = __fadd_rn(oz, __fmul_rn(t, dz));
  float o0 = __fdiv_rn(__fmul_rn(sx, ox), oz);
  float o1 = __fdiv_rn(__fmul_rn(sy, oy), oz);
  float o2 = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(2.0f, near), oz));
  float d0 = __fmul_rn(sx, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(ox, oz)));
  float d1 = __fmul_rn(sy, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(oy, oz)));
  float d2 = __fdiv_rn(__fmul_rn(-2.0f, near), oz);
  ox = o0; oy = o1; oz = o2; dx = d0; dy = d1; dz = d2;
}

__global__ void gen_rays_kernel(const float* __restrict__ poses, int n_views, int pose_rows, int H,
                                int W, float focal, const int64_t* __restrict__ pixel_ids,
                                int64_t first_id, int64_t n_rays, int ndc, float near, float sx,
                                float sy, const float* __restrict__ images,
                                float* __restrict__ rays_o, float* __restrict__ rays_d,
                                float* __restrict__ rgb_gt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  int64_t p = pixel_ids ? pixel_ids[i] : first_id + i;
  int64_t hw = (int64_t)H * W;
  int v = (int)(p / hw);
  int rem = (int)(p - (int64_t)v * hw);
  int h = rem / W, w = rem - h * W;
  const float* P = poses + (int64_t)v * pose_rows * 4;
  float x = __fdiv_rn(__fsub_rn((float)w, W * 0.5f), focal);
  float y = -__fdiv_rn(__fsub_rn((float)h, H * 0.5f), focal);
  float z = -1.0f;
  float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
  float d[3], o[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    d[a] = __fadd_rn(__fadd_rn(__fmul_rn(x, P[a * 4 + 0]), __fmul_rn(y, P[a * 4 + 1])),
                     __fmul_rn(z, P[a * 4 + 2]));
    o[a] = P[a * 4 + 3];
  }
  if (ndc) ndc_warp(o[0], o[1], o[2], d[0], d[1], d[2], near, sx, sy);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    rays_o[i * 3 + a] = o[a];
    rays_d[i * 3 + a] = d[a];
  }
  if (images && rgb_gt) {
#pragma unroll
    for (int a = 0; a < 3; ++a) rgb_gt[i * 3 + a] = images[p * 3 + a];
  }
}

__global__ void to_ndc_kernel(const float* __restrict__ ro, const float* __restrict__ rd, int64_t n,
                              float near, float sx, float sy, float* __restrict__ no,
                              float* __restrict__ nd) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ox = ro[i * 3], oy = ro[i * 3 + 1], oz = ro[i * 3 + 2];
  float dx = rd[i * 3], dy = rd[i * 3 + 1], dz = rd[i * 3 + 2];
  ndc_warp(ox, oy, oz, dx, dy, dz, near, sx, sy);
  no[i * 3] = ox; no[i * 3 + 1] = oy; no[i * 3 + 2] = oz;
  nd[i * 3] = dx; nd[i * 3 + 1] = dy; nd[i * 3 + 2] = dz;
}

// ---- counter-based uniforms (oracle/sampling.py:rng_uniform) -------------
// u(seed, i) in [0,1) with 24 random bits: two integer finalisers (murmur3 fmix32, then
// lowbias32) over the element index, keyed by the two halves of the seed.  Stateless, so the
// samplers draw their jitter in registers instead of reading a torch.rand buffer
// (reference: torch.rand at src/render/rendering.py — any i.i.d. U[0,1) stream is equivalent).
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t i) {
  uint32_t h = (uint32_t)i * 0x9E3779B1u + (uint32_t)(i >> 32) * 0x85EBCA77u + (uint32_t)seed;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  h += (uint32_t)(seed >> 32);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return (float)(h >> 8) * 0x1p-24f;
}

__global__ void rng_uniform_kernel(int64_t n, uint64_t seed, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rng_uniform(seed, (uint64_t)i);
}

// ---- stratified (oracle/sampling.py:stratified) ----
__device__ __forceinline__ float strat_z(int i, int S, float near, float far) {
  float t = (S > 1) ? __fdiv_rn((float)i, (float)(S - 1)) : 0.0f;
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}

// HBM bound (24 B/ray + 12 B/sample): the per-stratum bounds depend on the sample index only and
// are tabulated once per block in shared memory (the IEEE divisions of linspace are the expensive
// part); every point is computed ONCE and written as t_starts[i] and t_ends[i-1].
constexpr int kStratMaxS = 2048;
__global__ void __launch_bounds__(256)
stratified_kernel(int64_t n_rays, int S, float near, float far, const float* __restrict__ u, bool seeded,
                  uint64_t seed, float* __restrict__ ts, float* __restrict__ te, int rays_per_block) {
  __shared__ float lo[kStratMaxS], hi[kStratMaxS];
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float z = strat_z(i, S, near, far);
    lo[i] = (i == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, strat_z(i - 1, S, near, far)));
    hi[i] = (i == S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(strat_z(i + 1, S, near, far), z));
    if (!u && !seeded) lo[i] = hi[i] = z;  // deterministic: the points themselves
  }
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rays_per_block;
  const int64_t r1 = (r0 + rays_per_block < n_rays) ? r0 + rays_per_block : n_rays;
  const int64_t base = r0 * S;
  const int n = (int)((r1 - r0) * S);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e % S;
    const float l = lo[i];
    float p = l;
    if (u || seeded) {
      const float uv = seeded ? rng_uniform(seed, (uint64_t)(base + e)) : __ldg(u + base + e);
      p = __fadd_rn(l, __fmul_rn(__fsub_rn(hi[i], l), uv));
    }
    ts[base + e] = p;
    if (i > 0) te[base + e - 1] = p;
    if (i == S - 1) te[base + e] = far;
  }
}

// ---- sample_pdf (Appendix B2; oracle/sampling.py:sample_pdf) --------------
// One warp per ray.  CDF in "warp-tree order": lane l owns E consecutive
// weights; lane-local sums sequential, total by xor-butterfly, cross-lane
// prefix by Kogge-Stone.
constexpr int kPdfMaxE = 8;  // n_coarse - 2 <= 256
constexpr int kPdfWarps = 4;

// E = weights per lane = ceil((Sc - 2) / 32), a template parameter so that the unrolled CDF
// loops (each slot carries an IEEE division) contain no predicated-off slots.
template <int E>
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_kernel(int64_t n_rays, int Sc, int Sf, const float* __restrict__ z_coarse,
                  const float* __restrict__ w_coarse, const float* __restrict__ u, bool seeded,
                  uint64_t seed, float far, float* __restrict__ samples, int32_t* __restrict__ inds_out,
                  int32_t* __restrict__ perm_out, float* __restrict__ ts, float* __restrict__ te) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kPdfWarps + warp;
  const int nb = Sc - 1, nw = Sc - 2, ntot = Sc + Sf;
  int P2 = 1;                           // fine samples padded to a power of two for the sort
  while (P2 < Sf) P2 <<= 1;
  const int per_warp = 2 * nb + 2 * ntot + 2 * P2;
  float* cdf = smem + warp * per_warp;  // [nb]
  float* bins = cdf + nb;               // [nb]
  float* cat = bins + nb;               // [ntot]
  float* sorted = cat + ntot;           // [ntot]
  float* fk = sorted + ntot;            // [P2] fine samples, sorted in place
  int* fi = reinterpret_cast<int*>(fk + P2);  // [P2] their original indices
  if (r >= n_rays) return;
  const float* zc = z_coarse + r * Sc;
  const float* wc = w_coarse + r * Sc;

  for (int j = lane; j < nb; j += 32) bins[j] = __fmul_rn(0.5f, __fadd_rn(zc[j + 1], zc[j]));
  for (int j = lane; j < Sc; j += 32) cat[j] = zc[j];

  float wp[E];
  float s = 0.0f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    int j = lane * E + e;
    wp[e] = (j < nw) ? __fadd_rn(fmaxf(wc[j + 1], 0.0f), 1e-5f) : 0.0f;
    s = (e == 0) ? wp[0] : __fadd_rn(s, wp[e]);
  }
  float tot = s;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) tot = __fadd_rn(tot, __shfl_xor_sync(0xffffffffu, tot, m));
  float loc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float pdf = __fdiv_rn(wp[e], tot);
    loc[e] = (e == 0) ? pdf : __fadd_rn(loc[e > 0 ? e - 1 : 0], pdf);
  }
  float T = loc[E - 1];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float up = __shfl_up_sync(0xffffffffu, T, d);
    if (lane >= d) T = __fadd_rn(T, up);
  }
  float excl = __shfl_up_sync(0xffffffffu, T, 1);
  if (lane == 0) {
    excl = 0.0f;
    cdf[0] = 0.0f;
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    int j = lane * E + e;
    if (j < nw) cdf[j + 1] = (e == E - 1) ? T : __fadd_rn(excl, loc[e]);
  }
  __syncwarp();

  for (int k = lane; k < Sf; k += 32) {
    float uk = seeded ? rng_uniform(seed, (uint64_t)(r * Sf + k))
               : u    ? u[r * Sf + k]
                      : (Sf > 1 ? __fdiv_rn((float)k, (float)(Sf - 1)) : 0.0f);
    // searchsorted(right=True): number of cdf entries <= uk (cdf is increasing)
    int lo = 0, hi = nb;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uk) lo = mid + 1; else hi = mid;
    }
    int ind = lo;
    int below = max(ind - 1, 0), above = min(ind, nb - 1);
    float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.0f;
    float t = __fdiv_rn(__fsub_rn(uk, c0), denom);
    float smp = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    cat[Sc + k] = smp;
    if (samples) samples[r * Sf + k] = smp;
    if (inds_out) inds_out[r * Sf + k] = ind;
  }
  __syncwarp();

  // stable sort of cat[0..ntot) = merge of the (non-decreasing) coarse points with the fine
  // samples.  rank(i) = #{j: cat[j] < v or (== and j < i)}, computed as
  //   * fine samples: bitonic sort by (value, index) in shared memory (O(Sf log^2 Sf));
  //   * coarse j -> j + #{fine < z_j};  fine at sorted position q -> q + #{coarse <= v}
  //     (a coarse point precedes an equal fine sample: its index is smaller) by binary search.
  bool in_order = true;  // deterministic u (evaluation / rendering): the inverse CDF is monotone,
  for (int q = lane; q < P2; q += 32) {  // the fine samples arrive sorted and the network is skipped
    const float v = (q < Sf) ? cat[Sc + q] : INFINITY;
    fk[q] = v;
    fi[q] = q;
    if (q > 0 && q < Sf && cat[Sc + q - 1] > v) in_order = false;
  }
  in_order = __all_sync(0xffffffffu, in_order);
  __syncwarp();
  for (int k = 2; k <= P2 && !in_order; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P2 >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const float a = fk[i], b = fk[p];
        const int ia = fi[i], ib = fi[p];
        const bool a_gt_b = (a > b) || (a == b && ia > ib);
        if (a_gt_b == ((i & k) == 0)) {
          fk[i] = b; fk[p] = a;
          fi[i] = ib; fi[p] = ia;
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < Sc; j += 32) {
    const float v = cat[j];
    int lo = 0, hi = Sf;  // first fine position with value >= v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (fk[mid] < v) lo = mid + 1; else hi = mid;
    }
    sorted[j + lo] = v;
    if (perm_out) perm_out[r * ntot + j + lo] = j;
  }
  for (int q = lane; q < Sf; q += 32) {
    const float v = fk[q];
    int lo = 0, hi = Sc;  // first coarse position with value > v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cat[mid] <= v) lo = mid + 1; else hi = mid;
    }
    sorted[q + lo] = v;
    if (perm_out) perm_out[r * ntot + q + lo] = Sc + fi[q];
  }
  __syncwarp();
  for (int k = lane; k < ntot; k += 32) {
    ts[r * ntot + k] = sorted[k];
    te[r * ntot + k] = (k == ntot - 1) ? far : sorted[k + 1];
  }
}

}  // namespace

extern "C" int fsnerf_gen_rays(const float* poses, int n_views, int pose_rows, int H, int W,
                               float focal, const int64_t* pixel_ids, int64_t first_id,
                               int64_t n_rays, int ndc, float ndc_near, float ndc_sx, float ndc_sy,
                               const float* images, float* rays_o, float* rays_d, float* rgb_gt,
                               void* stream) {
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(poses && rays_o && rays_d, "gen_rays: null pointer");
  FS_REQUIRE(pose_rows == 3 || pose_rows == 4, "gen_rays: pose_rows must be 3 or 4");
  FS_REQUIRE(H > 0 && W > 0 && n_views > 0 && n_rays >= 0, "gen_rays: bad sizes");
  if (n_rays == 0) return FSNERF_OK;
  const int threads = 256;
  int64_t blocks = (n_rays + threads - 1) / threads;
  FsProfScope prof_("gen_rays", stream);
  gen_rays_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      poses, n_views, pose_rows, H, W, focal, pixel_ids, first_id, n_rays, ndc, ndc_near, ndc_sx,
      ndc_sy, images, rays_o, rays_d, rgb_gt);
  return fsnerf_check_launch("gen_rays");
}

extern "C" int fsnerf_to_ndc(const float* rays_o, const float* rays_d, int64_t n_rays, float near,
                             float sx, float sy, float* ndc_o, float* ndc_d, void* stream) {
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(rays_o && rays_d && ndc_o && ndc_d, "to_ndc: null pointer");
  const int threads = 256;
  int64_t blocks = (n_rays + threads - 1) / threads;
  FsProfScope prof_("to_ndc", stream);
  to_ndc_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays,
                                                                        near, sx, sy, ndc_o, ndc_d);
  return fsnerf_check_launch("to_ndc");
}

// ---------------------------------------------------------------- standalone encoding
// M.PositionalEncoder.forward (src/core/models.py:43-50) for callers that want the encoding
// itself: out[p] = [x, sin(f0 x), cos(f0 x), ..., sin(f_{L-1} x), cos(f_{L-1} x)] (x mask).
// Inside the MLP the encoding is never materialised (mlp_encode.cuh); this kernel is HBM
// bound: 4 d in, 4 d (1 + 2L) out per point.  One thread per (point, frequency slot).
namespace {
__global__ void encode_kernel(int64_t n, int d_in, int n_freqs, const float* __restrict__ freqs,
                              const float* __restrict__ mask, const float* __restrict__ x,
                              float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slots = 1 + n_freqs;
  if (i >= n * slots) return;
  const int64_t p = i / slots;
  const int k = (int)(i - p * slots);  // 0: identity, 1 + j: frequency j
  const int d_out = d_in * (1 + 2 * n_freqs);
  float* o = out + p * d_out;
  for (int a = 0; a < d_in; ++a) {
    const float v = x[p * d_in + a];
    if (k == 0) {
      o[a] = mask ? v * mask[a] : v;
    } else {
      float sn, cs;
      sincosf(v * freqs[k - 1], &sn, &cs);
      const int c = d_in + 2 * d_in * (k - 1) + a;
      o[c] = mask ? sn * mask[c] : sn;
      o[c + d_in] = mask ? cs * mask[c + d_in] : cs;
    }
  }
}
}  // namespace

extern "C" int fsnerf_encode(int64_t n_points, int d_input, int n_freqs, const float* freqs,
                             const float* mask, const float* x, float* out, void* stream) {
  if (n_points == 0) return FSNERF_OK;
  FS_REQUIRE(x && out && (freqs || n_freqs == 0), "encode: null pointer");
  FS_REQUIRE(d_input >= 1 && n_freqs >= 0 && n_freqs <= 64, "encode: bad sizes");
  const int64_t total = n_points * (1 + n_freqs);
  FsProfScope prof_("encode", stream);
  encode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_points, d_input, n_freqs, freqs,
                                                                                   mask, x, out);
  return fsnerf_check_launch("encode");
}

static int sample_stratified_impl(int64_t n_rays, int n_samples, float near, float far, const float* u,
                                  bool seeded, uint64_t seed, float* t_starts, float* t_ends, void* stream) {
  FS_REQUIRE(n_samples >= 1 && n_rays >= 0, "sample_stratified: bad sizes");
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(t_starts && t_ends, "sample_stratified: null output");
  FS_REQUIRE(n_samples <= kStratMaxS, "sample_stratified: n_samples must be <= 2048");
  FsProfScope prof_("sample_stratified", stream);
  int rays_per_block = 4096 / n_samples;  // ~4 K samples per block, at least one ray
  if (rays_per_block < 1) rays_per_block = 1;
  const int64_t blocks = (n_rays + rays_per_block - 1) / rays_per_block;
  stratified_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      n_rays, n_samples, near, far, u, seeded, seed, t_starts, t_ends, rays_per_block);
  return fsnerf_check_launch("sample_stratified");
}

extern "C" int fsnerf_sample_stratified(int64_t n_rays, int n_samples, float near, float far,
                                        const float* u, float* t_starts, float* t_ends,
                                        void* stream) {
  return sample_stratified_impl(n_rays, n_samples, near, far, u, false, 0, t_starts, t_ends, stream);
}

extern "C" int fsnerf_sample_stratified_seeded(int64_t n_rays, int n_samples, float near, float far,
                                               uint64_t seed, float* t_starts, float* t_ends,
                                               void* stream) {
  return sample_stratified_impl(n_rays, n_samples, near, far, nullptr, true, seed, t_starts, t_ends, stream);
}

extern "C" int fsnerf_rng_uniform(int64_t n, uint64_t seed, float* out, void* stream) {
  FS_REQUIRE(n >= 0, "rng_uniform: bad size");
  if (n == 0) return FSNERF_OK;
  FS_REQUIRE(out, "rng_uniform: null output");
  FsProfScope prof_("rng_uniform", stream);
  rng_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, seed, out);
  return fsnerf_check_launch("rng_uniform");
}

static int sample_pdf_impl(int64_t n_rays, int n_coarse, int n_fine, const float* z_coarse,
                           const float* w_coarse, const float* u, bool seeded, uint64_t seed, float far,
                           float* samples, int32_t* inds, int32_t* perm, float* t_starts, float* t_ends,
                           void* stream) {
  if (n_rays == 0) return FSNERF_OK;
  FS_REQUIRE(z_coarse && w_coarse && t_starts && t_ends, "sample_pdf: null pointer");
  FS_REQUIRE(n_coarse >= 3 && n_coarse - 2 <= 32 * kPdfMaxE, "sample_pdf: n_coarse must be in [3,258]");
  FS_REQUIRE(n_fine >= 1 && n_fine <= 1024, "sample_pdf: n_fine must be in [1,1024]");
  if (n_rays == 0) return FSNERF_OK;
  int p2 = 1;
  while (p2 < n_fine) p2 <<= 1;
  size_t smem = (size_t)kPdfWarps * (2 * (n_coarse - 1) + 2 * (n_coarse + n_fine) + 2 * p2) * sizeof(float);
  FS_REQUIRE(smem <= 48 * 1024, "sample_pdf: n_coarse+n_fine too large for shared memory");
  int64_t blocks = (n_rays + kPdfWarps - 1) / kPdfWarps;
  FsProfScope prof_("sample_pdf", stream);
#define LAUNCH_PDF(E_)                                                                          \
  sample_pdf_kernel<E_><<<(unsigned)blocks, kPdfWarps * 32, smem, (cudaStream_t)stream>>>(     \
      n_rays, n_coarse, n_fine, z_coarse, w_coarse, u, seeded, seed, far, samples, inds, perm, t_starts, t_ends)
  switch ((n_coarse - 2 + 31) / 32) {
    case 1: LAUNCH_PDF(1); break;
    case 2: LAUNCH_PDF(2); break;
    case 3: LAUNCH_PDF(3); break;
    case 4: LAUNCH_PDF(4); break;
    case 5: LAUNCH_PDF(5); break;
    case 6: LAUNCH_PDF(6); break;
    case 7: LAUNCH_PDF(7); break;
    default: LAUNCH_PDF(8); break;
  }
#undef LAUNCH_PDF
  return fsnerf_check_launch("sample_pdf");
}

extern "C" int fsnerf_sample_pdf(int64_t n_rays, int n_coarse, int n_fine, const float* z_coarse,
                                 const float* w_coarse, const float* u, float far, float* samples,
                                 int32_t* inds, int32_t* perm, float* t_starts, float* t_ends,
                                 void* stream) {
  return sample_pdf_impl(n_rays, n_coarse, n_fine, z_coarse, w_coarse, u, false, 0, far, samples, inds, perm,
                         t_starts, t_ends, stream);
}

extern "C" int fsnerf_sample_pdf_seeded(int64_t n_rays, int n_coarse, int n_fine, const float* z_coarse,
                                        const float* w_coarse, uint64_t seed, float far, float* samples,
                                        int32_t* inds, int32_t* perm, float* t_starts, float* t_ends,
                                        void* stream) {
  return sample_pdf_impl(n_rays, n_coarse, n_fine, z_coarse, w_coarse, nullptr, true, seed, far, samples, inds,
                         perm, t_starts, t_ends, stream);
}
