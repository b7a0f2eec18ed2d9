// Frequency-masked positional encoding of one sample row into a [128 x 64] bf16
// SWIZZLE_128B operand tile (kernel (2), fused into the MLP's A-operand staging).
#pragma once
#include "common.cuh"
#include "mlp_common.cuh"

namespace fs {

// out of line on purpose: sincosf's argument reduction is ~150 instructions and the fused
// MLP kernels must keep their hot loops instruction-cache resident
static __device__ __noinline__ void sincos_acc(float x, float* sn, float* cs) { sincosf(x, sn, cs); }

// sin/cos encoding of v[3] -> 64 bf16 channels (zero padded) into row `row` of a
// [128 x 64] SW128 tile at smem address `tile`.
// reference: src/core/models.py:43-50 (channel order x, sin(f0 x), cos(f0 x), ...)
static __device__ __noinline__ void encode_row(const float v[3], int n_freqs, const float* freqs, bool pow2,
                                           const float* __restrict__ mask, uint32_t tile, int row) {
  float ch[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) ch[i] = 0.f;
  ch[0] = v[0]; ch[1] = v[1]; ch[2] = v[2];
  if (pow2) {
    // f_k = 2^k (log_space, the reference default): an accurate sincosf every 4th octave,
    // exact double-angle steps (sin 2a = 2 s c, cos 2a = 1 - 2 s^2) in between.  The error
    // doubles per step, so it stays <= ~8 ulp (1e-6) — invisible after the bf16 rounding.
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float sn = 0.f, cs = 1.f;
#pragma unroll
      for (int k = 0; k < kMaxFreqs; ++k) {
        if (k < n_freqs) {
          if ((k & 3) == 0) {
            sincos_acc(v[a] * freqs[k], &sn, &cs);
          } else {
            const float s2 = 2.0f * sn * cs;
            cs = fmaf(-2.0f * sn, sn, 1.0f);
            sn = s2;
          }
          ch[3 + 6 * k + a] = sn;
          ch[3 + 6 * k + 3 + a] = cs;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < kMaxFreqs; ++k) {
      if (k < n_freqs) {
        float f = freqs[k];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float sn, cs;
          sincos_acc(v[a] * f, &sn, &cs);
          ch[3 + 6 * k + a] = sn;
          ch[3 + 6 * k + 3 + a] = cs;
        }
      }
    }
  }
  if (mask) {
    const int d = 3 + 6 * n_freqs;
#pragma unroll
    for (int i = 0; i < 63; ++i)
      if (i < d) ch[i] *= __ldg(mask + i);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    st_shared_v4(tile + sw128_off(row, j), pack_bf16x2(ch[8 * j], ch[8 * j + 1]),
                 pack_bf16x2(ch[8 * j + 2], ch[8 * j + 3]), pack_bf16x2(ch[8 * j + 4], ch[8 * j + 5]),
                 pack_bf16x2(ch[8 * j + 6], ch[8 * j + 7]));
  }
}

}  // namespace fs
