// rays.cu — camera rays of a pinhole view, written directly as the [N, 8 | 11] ray batch of the hot path.
//   replaces DS_NeRF/run_nerf_helpers.py:249-260 (get_rays) and the ray-batch assembly of render(), run.py:1171-1207
//   (patch crop, viewdir normalisation, near / far columns, torch.cat).
// Bit-exact against the reference on CPU (verified in the build container against the unmodified get_rays / torch.norm):
//   dirs   = ((j - W/2) / focal, -((i - H/2) / focal), -1)             IEEE division, j / i exact integers
//   rays_d = (dirs.x * R[k][0] + dirs.y * R[k][1]) + dirs.z * R[k][2]    no FMA contraction (torch.sum over 3 products)
//   |v|    = sqrt(fma(v.z, v.z, fma(v.y, v.y, v.x * v.x)))              torch.norm's accumulation order
//   ndc_rays (run_nerf_helpers.py:283-300) fused behind a flag; mvip_rays_pack does the same assembly for given rays.
#include "common.cuh"

namespace {
// NDC warp of one ray (run_nerf_helpers.py:283-300), every op rounded to fp32 in the reference's order:
//   t = -(near + o.z) / d.z;  o' = o + t*d (mul, then add);  sx = (float)(-1/(W/(2 focal))) computed in double by the host;
//   o_ndc = ((sx*o'.x)/o'.z, (sy*o'.y)/o'.z, 1 + (1/o'.z)*2near);  d_ndc = (sx*(d.x/d.z - o'.x/o'.z), sy*(...), (1/o'.z)*(-2near))
//   (`c / tensor` is reciprocal(tensor) * c in torch, Tensor.__rtruediv__).
struct Ndc { int on; float near, sx, sy, two_near, neg_two_near; };

__device__ __forceinline__ void emit_ray(float* __restrict__ o, const float ro[3], const float rd[3], const float v[3],
                                         float near, float far, int use_viewdirs, const Ndc& ndc) {
  if (ndc.on) {
    const float t = __fdiv_rn(-__fadd_rn(ndc.near, ro[2]), rd[2]);
    const float ox = __fadd_rn(ro[0], __fmul_rn(t, rd[0]));
    const float oy = __fadd_rn(ro[1], __fmul_rn(t, rd[1]));
    const float oz = __fadd_rn(ro[2], __fmul_rn(t, rd[2]));
    const float rz = __fdiv_rn(1.f, oz);
    const float qx = __fdiv_rn(ox, oz), qy = __fdiv_rn(oy, oz);
    o[0] = __fdiv_rn(__fmul_rn(ndc.sx, ox), oz);
    o[1] = __fdiv_rn(__fmul_rn(ndc.sy, oy), oz);
    o[2] = __fadd_rn(1.f, __fmul_rn(rz, ndc.two_near));
    o[3] = __fmul_rn(ndc.sx, __fsub_rn(__fdiv_rn(rd[0], rd[2]), qx));
    o[4] = __fmul_rn(ndc.sy, __fsub_rn(__fdiv_rn(rd[1], rd[2]), qy));
    o[5] = __fmul_rn(rz, ndc.neg_two_near);
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ro[k]; o[3 + k] = rd[k]; }
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

// block-wide: rows of rays [blk, blk + 256) staged in shared memory -> global, consecutive floats per consecutive thread
__device__ __forceinline__ void flush_rows(const float* rows, float* __restrict__ out, int64_t blk, int64_t n, int stride) {
  __syncthreads();
  const int64_t left = n - blk;
  const int total = (int)(left < 256 ? left : 256) * stride;
  float* dst = out + blk * stride;
  for (int e = threadIdx.x; e < total; e += 256) dst[e] = rows[e];
  __syncthreads();
}

__global__ void __launch_bounds__(256) rays_kernel(const float* __restrict__ c2w, const float* __restrict__ c2w_static, int H, int W,
                                                   float focal, float near, float far, int i0, int j0, int h, int w,
                                                   int use_viewdirs, Ndc ndc, float* __restrict__ out) {
  __shared__ float pose[24];
  if (threadIdx.x < 12) {
    pose[threadIdx.x] = c2w[threadIdx.x];
    pose[12 + threadIdx.x] = c2w_static ? c2w_static[threadIdx.x] : c2w[threadIdx.x];
  }
  __syncthreads();
  __shared__ float rows[256 * 11];   // rows are staged here and leave as consecutive 4-byte stores (a lane per float, not per ray)
  const int stride = use_viewdirs ? 11 : 8;
  const int64_t n = (int64_t)h * w;
  for (int64_t blk = (int64_t)blockIdx.x * blockDim.x; blk < n; blk += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = blk + threadIdx.x;
    if (idx < n) {
    const int i = i0 + (int)(idx / w), j = j0 + (int)(idx % w);
    const float dx = __fdiv_rn(__fsub_rn((float)j, __fmul_rn((float)W, .5f)), focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)i, __fmul_rn((float)H, .5f)), focal);
    const float dz = -1.f;
    float ro[3], rd[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // rays_o / rays_d from the (possibly static) camera
      const float* R = pose + 12 + 4 * k;
      ro[k] = R[3];
      rd[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[0]), __fmul_rn(dy, R[1])), __fmul_rn(dz, R[2]));
      const float* Rv = pose + 4 * k;
      v[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, Rv[0]), __fmul_rn(dy, Rv[1])), __fmul_rn(dz, Rv[2]));
    }
    emit_ray(rows + threadIdx.x * stride, ro, rd, v, near, far, use_viewdirs, ndc);
    }
    flush_rows(rows, out, blk, n, stride);
  }
}

// the `rays=` entry of render() (run.py:1176-1207): given origins / directions -> the same packed batch
__global__ void __launch_bounds__(256) rays_pack_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ view_d, int64_t n, float near, float far,
                                                        int use_viewdirs, Ndc ndc, float* __restrict__ out) {
  __shared__ float rows[256 * 11];
  const int stride = use_viewdirs ? 11 : 8;
  for (int64_t blk = (int64_t)blockIdx.x * blockDim.x; blk < n; blk += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = blk + threadIdx.x;
    if (idx < n) {
      float ro[3], rd[3], v[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ro[k] = rays_o[idx * 3 + k];
        rd[k] = rays_d[idx * 3 + k];
        v[k] = view_d ? view_d[idx * 3 + k] : rd[k];
      }
      emit_ray(rows + threadIdx.x * stride, ro, rd, v, near, far, use_viewdirs, ndc);
    }
    flush_rows(rows, out, blk, n, stride);
  }
}

Ndc make_ndc(int on, int H, int W, double focal, double ndc_near) {
  Ndc c{};
  c.on = on ? 1 : 0;
  if (on) {   // python-double scalars, cast to fp32 when they meet a tensor
    c.near = (float)ndc_near;
    c.sx = (float)(-1. / ((double)W / (2. * focal)));
    c.sy = (float)(-1. / ((double)H / (2. * focal)));
    c.two_near = (float)(2. * ndc_near);
    c.neg_two_near = (float)(-2. * ndc_near);
  }
  return c;
}

unsigned grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)mvip_num_sms() * 16;
  return (unsigned)(blocks > cap ? cap : blocks);
}
}  // namespace

extern "C" int mvip_rays_from_pose_ndc(const float* c2w, const float* c2w_static, int H, int W, double focal, float near,
                                       float far, int i0, int j0, int h, int w, int use_viewdirs, int ndc, double ndc_near,
                                       float* out, void* stream) {
  MVIP_REQUIRE(H > 0 && W > 0 && h >= 0 && w >= 0 && i0 >= 0 && j0 >= 0 && i0 + h <= H && j0 + w <= W, MVIP_E_INVALID,
               "mvip_rays_from_pose: window [%d,+%d) x [%d,+%d) outside the %d x %d image", i0, h, j0, w, H, W);
  MVIP_REQUIRE(focal != 0., MVIP_E_INVALID, "mvip_rays_from_pose: focal == 0");
  const int64_t n = (int64_t)h * w;
  if (n == 0) return MVIP_OK;
  MVIP_REQUIRE(c2w && out, MVIP_E_INVALID, "mvip_rays_from_pose: null pointer");
  rays_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(c2w, c2w_static, H, W, (float)focal, near, far, i0, j0, h, w,
                                                             use_viewdirs ? 1 : 0, make_ndc(ndc, H, W, focal, ndc_near), out);
  MVIP_LAUNCH_OK("rays_kernel");
  return MVIP_OK;
}

extern "C" int mvip_rays_from_pose(const float* c2w, const float* c2w_static, int H, int W, float focal, float near,
                                   float far, int i0, int j0, int h, int w, int use_viewdirs, float* out, void* stream) {
  return mvip_rays_from_pose_ndc(c2w, c2w_static, H, W, (double)focal, near, far, i0, j0, h, w, use_viewdirs, 0, 0., out, stream);
}

extern "C" int mvip_rays_pack(const float* rays_o, const float* rays_d, const float* view_d, int64_t n, float near, float far,
                              int use_viewdirs, int ndc, int H, int W, double focal, double ndc_near, float* out, void* stream) {
  MVIP_REQUIRE(n >= 0, MVIP_E_INVALID, "mvip_rays_pack: n < 0");
  if (n == 0) return MVIP_OK;
  MVIP_REQUIRE(rays_o && rays_d && out, MVIP_E_INVALID, "mvip_rays_pack: null pointer");
  MVIP_REQUIRE(!ndc || (H > 0 && W > 0 && focal != 0.), MVIP_E_INVALID, "mvip_rays_pack: ndc needs H, W, focal");
  rays_pack_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, view_d, n, near, far, use_viewdirs ? 1 : 0,
                                                                  make_ndc(ndc, H, W, focal, ndc_near), out);
  MVIP_LAUNCH_OK("rays_pack_kernel");
  return MVIP_OK;
}
