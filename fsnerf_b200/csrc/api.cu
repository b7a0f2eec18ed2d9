// C-ABI plumbing (error reporting, device check) + the train-step arithmetic
// kernels (MSE gradient seed, fused Adam) of src/run-nerf.py:216-217,255-258,282-285.
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"
#include "../../include/fsnerf_b200.h"

static thread_local char g_err[512] = "";

void fsnerf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fsnerf_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fsnerf_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return FSNERF_ERR_CUDA;
  }
  return FSNERF_OK;
}

// ---------------------------------------------------------------- profiling
#include <string.h>
#include <vector>
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace

FsProfScope::FsProfScope(const char* name, void* st) : slot(-1), stream(st) {
  if (!g_prof_on) return;
  ProfRec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, (cudaStream_t)st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
FsProfScope::~FsProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, (cudaStream_t)stream);
}

extern "C" int fsnerf_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = on != 0;
  return FSNERF_OK;
}

extern "C" int fsnerf_profile_read(int max_kernels, char* names, float* total_ms, int* counts) {
  FS_REQUIRE(names && total_ms && counts && max_kernels > 0, "profile_read: null pointer");
  int n = 0;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    int k = 0;
    for (; k < n; ++k) if (strncmp(names + 32 * k, r.name, 31) == 0) break;
    if (k == n) {
      if (n == max_kernels) continue;
      strncpy(names + 32 * n, r.name, 31);
      names[32 * n + 31] = 0;
      total_ms[n] = 0.f; counts[n] = 0;
      ++n;
    }
    total_ms[k] += ms;
    counts[k] += 1;
  }
  return n;
}

static void* g_trace = nullptr;
void* fsnerf_debug_trace_ptr() { return g_trace; }
extern "C" int fsnerf_debug_set_trace(void* buf) {
  g_trace = buf;
  return FSNERF_OK;
}

extern "C" int fsnerf_version(void) { return 100; }
extern "C" const char* fsnerf_last_error(void) { return g_err; }

extern "C" int fsnerf_device_ok(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    fsnerf_set_error("device_ok: %s", cudaGetErrorString(e));
    return FSNERF_ERR_CUDA;
  }
  if (p.major != 10) {
    fsnerf_set_error("device_ok: device %d is sm_%d%d, this library is sm_100a only", dev, p.major,
                     p.minor);
    return FSNERF_ERR_UNSUPPORTED;
  }
  return FSNERF_OK;
}

namespace {

__global__ void mse_loss_grad_kernel(int64_t n, const float* __restrict__ rgb,
                                     const float* __restrict__ gt, float grad_scale,
                                     float* __restrict__ loss_sum, float* __restrict__ d_rgb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (i < n) {
    float d = rgb[i] - gt[i];
    sq = d * d;
    if (d_rgb) d_rgb[i] = grad_scale * 2.0f * d;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, m);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0 && loss_sum) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(loss_sum, s);
  }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False) single-tensor form:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, float step_size,
                            float b1, float b2, float eps, float inv_sqrt_bc2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = b1 * m[i] + (1.0f - b1) * gi;
  float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

}  // namespace

extern "C" int fsnerf_mse_loss_grad(int64_t n, const float* rgb, const float* gt, float grad_scale,
                                    float* loss_sum, float* d_rgb, void* stream) {
  FS_REQUIRE(rgb && gt, "mse_loss_grad: null pointer");
  if (n <= 0) return FSNERF_OK;
  int threads = 256;
  FsProfScope prof_("mse_loss_grad", stream);
  mse_loss_grad_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, rgb, gt, grad_scale, loss_sum, d_rgb);
  return fsnerf_check_launch("mse_loss_grad");
}

extern "C" int fsnerf_adam_step(int64_t n, float* params, const float* grads, float* m, float* v,
                                float lr, float beta1, float beta2, float eps, int step,
                                void* stream) {
  FS_REQUIRE(params && grads && m && v, "adam_step: null pointer");
  FS_REQUIRE(step >= 1, "adam_step: step counts from 1");
  if (n <= 0) return FSNERF_OK;
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  float step_size = (float)((double)lr / bc1);
  float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int threads = 256;
  FsProfScope prof_("adam", stream);
  adam_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, params, grads, m, v, step_size, beta1, beta2, eps, inv_sqrt_bc2);
  return fsnerf_check_launch("adam_step");
}
