// MLP backward (placeholder until the dgrad/wgrad kernels land).
#include "common.cuh"
#include "mlp_common.cuh"

extern "C" int64_t fsnerf_mlp_bwd_workspace_bytes(const fsnerf_net_cfg* cfg, int64_t n_samples) {
  (void)cfg; (void)n_samples;
  return 0;
}
extern "C" int fsnerf_mlp_backward(const fsnerf_net_cfg* cfg, const float* params,
                                   const void* packed, int64_t n_samples, const void* stash,
                                   const float* out, const float* d_out, int density_only,
                                   float* grads, void* workspace, void* stream) {
  fsnerf_set_error("mlp_backward: not implemented yet");
  return FSNERF_ERR_UNSUPPORTED;
}
