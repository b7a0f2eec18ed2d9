/* Plain-C client of the C ABI (include/fsnerf_b200.h): no Python, no torch — device buffers from
 * cudaMalloc, one stream, return codes + fsnerf_last_error().  Built and run by
 * tests/test_gpu_cabi_c.py:
 *   gcc -O2 -I include -I /usr/local/cuda/include tests/cabi_smoke.c -o cabi_smoke \
 *       -L fsnerf_b200 -lfsnerf_b200 -L /usr/local/cuda/lib64 -lcudart -lm
 * Checks, against straightforward C loops on the host:
 *   1. stratified sampling (deterministic) + alpha compositing forward (nerfacc semantics,
 *      src/render/rendering.py:89-96) of analytic (rgb, sigma) samples;
 *   2. ray generation for one pose (src/utils/utilities.py:36-82): centre pixel looks down -z;
 *   3. error path: a NULL output pointer returns a negative code and a message. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsnerf_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define FS(x) do { int r_ = (x); if (r_ != 0) { printf("%s -> %d: %s\n", #x, r_, fsnerf_last_error()); return 3; } } while (0)

int main(void) {
  printf("fsnerf_version = %d\n", fsnerf_version());
  FS(fsnerf_device_ok(0));
  CK(cudaSetDevice(0));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const int R = 257, S = 64;
  const float near = 2.0f, far = 6.0f;
  float *ts, *te, *raw, *rgb, *op, *dp, *w;
  CK(cudaMalloc((void**)&ts, sizeof(float) * R * S));
  CK(cudaMalloc((void**)&te, sizeof(float) * R * S));
  CK(cudaMalloc((void**)&raw, sizeof(float) * R * S * 4));
  CK(cudaMalloc((void**)&rgb, sizeof(float) * R * 3));
  CK(cudaMalloc((void**)&op, sizeof(float) * R));
  CK(cudaMalloc((void**)&dp, sizeof(float) * R));
  CK(cudaMalloc((void**)&w, sizeof(float) * R * S));
  FS(fsnerf_sample_stratified(R, S, near, far, NULL, ts, te, st));
  float* h_ts = (float*)malloc(sizeof(float) * R * S);
  float* h_te = (float*)malloc(sizeof(float) * R * S);
  float* h_raw = (float*)malloc(sizeof(float) * R * S * 4);
  CK(cudaMemcpyAsync(h_ts, ts, sizeof(float) * R * S, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_te, te, sizeof(float) * R * S, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < S; ++i) {  /* linspace(near, far, S), intervals [z_i, z_i+1], last one ends at far */
    float z = near + (far - near) * (float)i / (float)(S - 1);
    if (fabsf(h_ts[i] - z) > 1e-5f || fabsf(h_te[S - 1] - far) > 0) { printf("stratified mismatch at %d\n", i); return 4; }
  }
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      float* q = h_raw + 4 * ((size_t)r * S + s);
      q[0] = 0.5f + 0.5f * sinf(0.1f * s + r);
      q[1] = 0.25f;
      q[2] = (float)s / S;
      q[3] = 3.0f * expf(-0.5f * (s - 20.0f - 0.05f * r) * (s - 20.0f - 0.05f * r) / 16.0f) - 0.1f;  /* raw sigma may be < 0 */
    }
  CK(cudaMemcpyAsync(raw, h_raw, sizeof(float) * R * S * 4, cudaMemcpyHostToDevice, st));
  float bk[3] = {1.0f, 1.0f, 1.0f}, *d_bk;
  CK(cudaMalloc((void**)&d_bk, sizeof(bk)));
  CK(cudaMemcpyAsync(d_bk, bk, sizeof(bk), cudaMemcpyHostToDevice, st));
  FS(fsnerf_composite_forward(R, S, raw, ts, te, NULL, d_bk, 0, rgb, op, dp, w, NULL, NULL, st));
  float* h_rgb = (float*)malloc(sizeof(float) * R * 3);
  float* h_op = (float*)malloc(sizeof(float) * R);
  float* h_dp = (float*)malloc(sizeof(float) * R);
  CK(cudaMemcpyAsync(h_rgb, rgb, sizeof(float) * R * 3, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_op, op, sizeof(float) * R, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_dp, dp, sizeof(float) * R, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double worst = 0;
  for (int r = 0; r < R; ++r) {
    double acc = 0, c[3] = {0, 0, 0}, a = 0, d = 0;
    for (int s = 0; s < S; ++s) {
      const float* q = h_raw + 4 * ((size_t)r * S + s);
      double t0 = h_ts[(size_t)r * S + s], t1 = h_te[(size_t)r * S + s];
      double sd = (double)q[3] * (t1 - t0), T = exp(-acc), wt = T * (1.0 - exp(-sd));
      acc += sd;
      for (int k = 0; k < 3; ++k) c[k] += wt * q[k];
      a += wt;
      d += wt * 0.5 * (t0 + t1);
    }
    d /= (a > 1.1920929e-7 ? a : 1.1920929e-7);
    for (int k = 0; k < 3; ++k) {
      double e = fabs(c[k] + bk[k] * (1.0 - a) - h_rgb[3 * r + k]);
      if (e > worst) worst = e;
    }
    if (fabs(a - h_op[r]) > worst) worst = fabs(a - h_op[r]);
    if (fabs(d - h_dp[r]) > 1e-4) { printf("depth mismatch ray %d: %f vs %f\n", r, d, h_dp[r]); return 5; }
  }
  printf("composite: max |C ABI - host loop| = %.3g\n", worst);
  if (worst > 2e-5) return 6;

  /* 2. rays of a camera at (0,0,4) looking down -z */
  const int H = 100, W = 100;
  const float focal = 0.5f * W / tanf(0.5f * 0.6911112f);
  float pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 4, 0, 0, 0, 1}, *d_pose, *ro, *rd;
  CK(cudaMalloc((void**)&d_pose, sizeof(pose)));
  CK(cudaMalloc((void**)&ro, sizeof(float) * H * W * 3));
  CK(cudaMalloc((void**)&rd, sizeof(float) * H * W * 3));
  CK(cudaMemcpyAsync(d_pose, pose, sizeof(pose), cudaMemcpyHostToDevice, st));
  FS(fsnerf_gen_rays(d_pose, 1, 4, H, W, focal, NULL, 0, (int64_t)H * W, 0, 1.0f, 0.f, 0.f, NULL, ro, rd, NULL, st));
  float c_d[3], c_o[3], corner[3];
  CK(cudaMemcpyAsync(c_d, rd + 3 * (50 * W + 50), sizeof(c_d), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(c_o, ro + 3 * (50 * W + 50), sizeof(c_o), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(corner, rd, sizeof(corner), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  printf("centre ray d = (%g, %g, %g), o = (%g, %g, %g); corner d = (%.4f, %.4f, %.4f)\n", c_d[0], c_d[1], c_d[2],
         c_o[0], c_o[1], c_o[2], corner[0], corner[1], corner[2]);
  if (c_d[0] != 0.f || c_d[1] != 0.f || c_d[2] != -1.f || c_o[2] != 4.f) return 7;
  if (fabsf(corner[0] + 0.3208f) > 1e-3f || fabsf(corner[1] - 0.3208f) > 1e-3f || fabsf(corner[2] + 0.8912f) > 1e-3f) return 8;

  /* 3. error path */
  int rc = fsnerf_composite_forward(R, S, raw, ts, te, NULL, NULL, 0, NULL, op, dp, w, NULL, NULL, st);
  printf("null output -> %d (%s)\n", rc, fsnerf_last_error());
  if (rc >= 0 || strstr(fsnerf_last_error(), "null") == NULL) return 9;
  CK(cudaStreamSynchronize(st));
  printf("cabi_smoke: OK\n");
  return 0;
}
