// mlp_backward.cu — placeholder until the dgrad-chain / wgrad kernels land.
#include "mlp_common.cuh"
extern "C" {
size_t mvip_mlp_backward_workspace_bytes(int64_t n_points) { return (size_t)mlp::num_tiles(n_points) * mlp::kDzTileBytes; }
int mvip_mlp_backward(const void*, const float*, int64_t, const void*, void*, float* const*, int, void*) {
  mvip_set_error("mvip_mlp_backward: not implemented yet");
  return MVIP_E_UNSUPPORTED;
}
}
