// adam.cu — one-launch Adam step over a list of fp32 tensors.
//   replaces the per-tensor optimizer kernels behind optimizer.step() in the reference's training loop
//   (DS_NeRF/run.py:1003; optimizer built at run.py:1536-1537, learning-rate decay run.py:1031-1039 enters through `lr`).
// Semantics of torch.optim.Adam(params, lr, betas, eps) with weight_decay = 0, amsgrad = False, maximize = False:
//   m <- m + (1 - b1) (g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"
#include <math.h>

namespace {
constexpr int kMaxTensors = 64;

struct AdamArgs {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  int first[kMaxTensors + 1];   // prefix sums of the element counts
  int n;
  float step_size, beta1, beta2, one_m_beta1, one_m_beta2, inv_bc2_sqrt, eps;
  // CUDA-graph form: learning rate and 1-based step number live in device memory, so that one captured launch stays valid
  // while both change from replay to replay; the bias corrections are then formed in the kernel (in double, as on the host)
  const float* lr_dev;
  const long long* step_dev;
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  __shared__ float hyper[2];
  float step_size = a.step_size, inv_bc2_sqrt = a.inv_bc2_sqrt;
  if (a.step_dev != nullptr) {
    if (threadIdx.x == 0) {
      const double t = (double)*a.step_dev;
      hyper[0] = (float)((double)*a.lr_dev / (1.0 - pow((double)a.beta1, t)));
      hyper[1] = (float)(1.0 / sqrt(1.0 - pow((double)a.beta2, t)));
    }
    __syncthreads();
    step_size = hyper[0];
    inv_bc2_sqrt = hyper[1];
  }
  const int total = a.first[a.n];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int lo = 0, hi = a.n;                       // largest t with first[t] <= idx
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.first[mid] <= idx) lo = mid; else hi = mid;
    }
    const int e = idx - a.first[lo];
    const float g = a.g[lo][e];
    float m = a.m[lo][e], v = a.v[lo][e];
    m = m + a.one_m_beta1 * (g - m);
    v = a.beta2 * v + a.one_m_beta2 * g * g;
    a.m[lo][e] = m;
    a.v[lo][e] = v;
    const float denom = sqrtf(v) * inv_bc2_sqrt + a.eps;
    a.p[lo][e] -= step_size * (m / denom);
  }
}
}  // namespace

static int adam_launch(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                       const int64_t* sizes, int n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                       const float* lr_dev, const long long* step_dev, void* stream) {
  MVIP_REQUIRE(n_tensors >= 0 && step >= 1, MVIP_E_INVALID, "mvip_adam_step: bad n_tensors / step");
  MVIP_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && sizes), MVIP_E_INVALID, "mvip_adam_step: null array");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  for (int t0 = 0; t0 < n_tensors; t0 += kMaxTensors) {
    AdamArgs a;
    a.n = n_tensors - t0 < kMaxTensors ? n_tensors - t0 : kMaxTensors;
    a.first[0] = 0;
    for (int i = 0; i < a.n; ++i) {
      MVIP_REQUIRE(params[t0 + i] && grads[t0 + i] && exp_avg[t0 + i] && exp_avg_sq[t0 + i] && sizes[t0 + i] >= 0 &&
                       sizes[t0 + i] < (1ll << 30),
                   MVIP_E_INVALID, "mvip_adam_step: tensor %d: null pointer or bad size", t0 + i);
      a.p[i] = params[t0 + i]; a.g[i] = grads[t0 + i]; a.m[i] = exp_avg[t0 + i]; a.v[i] = exp_avg_sq[t0 + i];
      a.first[i + 1] = a.first[i] + (int)sizes[t0 + i];
      MVIP_REQUIRE(a.first[i + 1] >= a.first[i], MVIP_E_INVALID, "mvip_adam_step: more than 2^31 elements in one launch");
    }
    a.step_size = (float)((double)lr / bc1);
    a.beta1 = beta1; a.beta2 = beta2;
    a.one_m_beta1 = (float)(1.0 - (double)beta1); a.one_m_beta2 = (float)(1.0 - (double)beta2);   // as torch: 1 - beta in double
    a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    a.eps = eps;
    a.lr_dev = lr_dev;
    a.step_dev = step_dev;
    const int total = a.first[a.n];
    if (total == 0) continue;
    int blocks = (total + 255) / 256;
    const int cap = mvip_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    MVIP_LAUNCH_OK("adam_kernel");
  }
  return MVIP_OK;
}

extern "C" int mvip_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* sizes, int n_tensors, float lr, float beta1,
                              float beta2, float eps, int64_t step, void* stream) {
  return adam_launch(params, grads, exp_avg, exp_avg_sq, sizes, n_tensors, lr, beta1, beta2, eps, step, nullptr, nullptr, stream);
}

extern "C" int mvip_adam_step_dev(float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int64_t* sizes, int n_tensors, const float* lr_dev,
                                  float beta1, float beta2, float eps, const int64_t* step_dev, void* stream) {
  MVIP_REQUIRE(lr_dev && step_dev, MVIP_E_INVALID, "mvip_adam_step_dev: null lr / step pointer");
  return adam_launch(params, grads, exp_avg, exp_avg_sq, sizes, n_tensors, 0.f, beta1, beta2, eps, 1, lr_dev,
                     reinterpret_cast<const long long*>(step_dev), stream);
}
