// api.cu — error plumbing and device queries of the C ABI (include/mvip_nerf.h).
#include "mlp_common.cuh"
#include <stdlib.h>

static thread_local char g_err[512] = "";

void mvip_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mvip_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" {

int mvip_abi_version(void) { return MVIP_ABI_VERSION; }

const char* mvip_last_error(void) { return g_err; }

int mvip_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  MVIP_CUDA_OK(cudaGetDevice(&dev));
  MVIP_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MVIP_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

}  // extern "C"
