// rays.cu — camera rays of a pinhole view, written directly as the [N, 8 | 11] ray batch of the hot path.
//   replaces DS_NeRF/run_nerf_helpers.py:249-260 (get_rays) and the ray-batch assembly of render(), run.py:1171-1207
//   (patch crop, viewdir normalisation, near / far columns, torch.cat).
// Bit-exact against the reference on CPU (verified in the build container against the unmodified get_rays / torch.norm):
//   dirs   = ((j - W/2) / focal, -((i - H/2) / focal), -1)             IEEE division, j / i exact integers
//   rays_d = (dirs.x * R[k][0] + dirs.y * R[k][1]) + dirs.z * R[k][2]    no FMA contraction (torch.sum over 3 products)
//   |v|    = sqrt(fma(v.z, v.z, fma(v.y, v.y, v.x * v.x)))              torch.norm's accumulation order
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) rays_kernel(const float* __restrict__ c2w, const float* __restrict__ c2w_static, int H, int W,
                                                   float focal, float near, float far, int i0, int j0, int h, int w,
                                                   int use_viewdirs, float* __restrict__ out) {
  __shared__ float pose[24];
  if (threadIdx.x < 12) {
    pose[threadIdx.x] = c2w[threadIdx.x];
    pose[12 + threadIdx.x] = c2w_static ? c2w_static[threadIdx.x] : c2w[threadIdx.x];
  }
  __syncthreads();
  const int stride = use_viewdirs ? 11 : 8;
  const int64_t n = (int64_t)h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = i0 + (int)(idx / w), j = j0 + (int)(idx % w);
    const float dx = __fdiv_rn(__fsub_rn((float)j, __fmul_rn((float)W, .5f)), focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)i, __fmul_rn((float)H, .5f)), focal);
    const float dz = -1.f;
    float* o = out + idx * stride;
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // rays_o / rays_d from the (possibly static) camera
      const float* R = pose + 12 + 4 * k;
      o[k] = R[3];
      o[3 + k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[0]), __fmul_rn(dy, R[1])), __fmul_rn(dz, R[2]));
      const float* Rv = pose + 4 * k;
      v[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, Rv[0]), __fmul_rn(dy, Rv[1])), __fmul_rn(dz, Rv[2]));
    }
    o[6] = near;
    o[7] = far;
    if (use_viewdirs) {
      const float nrm = __fsqrt_rn(__fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0]))));
      o[8] = __fdiv_rn(v[0], nrm);
      o[9] = __fdiv_rn(v[1], nrm);
      o[10] = __fdiv_rn(v[2], nrm);
    }
  }
}
}  // namespace

extern "C" int mvip_rays_from_pose(const float* c2w, const float* c2w_static, int H, int W, float focal, float near,
                                   float far, int i0, int j0, int h, int w, int use_viewdirs, float* out, void* stream) {
  MVIP_REQUIRE(H > 0 && W > 0 && h >= 0 && w >= 0 && i0 >= 0 && j0 >= 0 && i0 + h <= H && j0 + w <= W, MVIP_E_INVALID,
               "mvip_rays_from_pose: window [%d,+%d) x [%d,+%d) outside the %d x %d image", i0, h, j0, w, H, W);
  MVIP_REQUIRE(focal != 0.f, MVIP_E_INVALID, "mvip_rays_from_pose: focal == 0");
  const int64_t n = (int64_t)h * w;
  if (n == 0) return MVIP_OK;
  MVIP_REQUIRE(c2w && out, MVIP_E_INVALID, "mvip_rays_from_pose: null pointer");
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)mvip_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  rays_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(c2w, c2w_static, H, W, focal, near, far, i0, j0, h, w,
                                                                  use_viewdirs ? 1 : 0, out);
  MVIP_LAUNCH_OK("rays_kernel");
  return MVIP_OK;
}
