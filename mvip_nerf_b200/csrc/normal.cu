// normal.cu — depth map -> least-squares plane normal map, forward and backward.
//   replaces DS_NeRF/run.py:1909-1922 (depth2xyz_torch) + :1924-1940 (depth2normal_geo)
//
// The reference unfolds a zero-padded k x k window per pixel (11.5 KB/pixel at k=31) and solves
// n = (A^T A)^-1 A^T 1.  Zero rows contribute nothing, so with a_i = (x,y,z) of pixel i:
//   M_p = sum_{i in win(p)} a_i a_i^T (6 unique),  s_p = sum a_i,  n_p = M_p^-1 s_p        (SURVEY §8a'-4)
// i.e. a 9-channel separable box filter + a 3x3 solve.  xyz is formed in fp32 exactly as the
// reference does; the box sums and the solve run in fp64 (M is ill-conditioned, cond ~1e4), which
// keeps us inside the reference's own fp32 noise.  Backward: q = M^-1 g, then
//   dL/da_i = sum_{p: i in win(p)} q_p - (sum_p q_p n_p^T + n_p q_p^T) a_i   -> the same box filter.
#include "common.cuh"

namespace {

struct Cam {
  float fx, fy, cx, cy;
  const float* xyz;   // non-null: take (x,y,z) from three [H,W] planes instead of back-projecting the depth
  size_t plane;
};

__device__ __forceinline__ void xyz_of(const float* __restrict__ depth, int W, int h, int w, const Cam& c, float& x,
                                       float& y, float& z) {
  if (c.xyz) {
    const size_t o = (size_t)h * W + w;
    x = __ldg(c.xyz + o); y = __ldg(c.xyz + c.plane + o); z = __ldg(c.xyz + 2 * c.plane + o);
    return;
  }
  z = __ldg(depth + (size_t)h * W + w);
  x = __fdiv_rn(__fmul_rn((float)w - c.cx, z), c.fx);  // (w-cx)*z/fx   run.py:1918
  y = __fdiv_rn(__fmul_rn((float)h - c.cy, z), c.fy);  // (h-cy)*z/fy   run.py:1919
}

// pass A: horizontal window sums of the 9 moment channels of a -> ws[9][H][W] (fp64)
__global__ void moments_h_kernel(const float* __restrict__ depth, int H, int W, Cam cam, int r, double* __restrict__ ws) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int h = blockIdx.y;
  if (w >= W) return;
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int lo = max(0, w - r), hi = min(W - 1, w + r);
  for (int j = lo; j <= hi; ++j) {
    float xf, yf, zf;
    xyz_of(depth, W, h, j, cam, xf, yf, zf);
    double x = xf, y = yf, z = zf;
    m[0] += x * x; m[1] += x * y; m[2] += x * z; m[3] += y * y; m[4] += y * z; m[5] += z * z;
    m[6] += x; m[7] += y; m[8] += z;
  }
  size_t plane = (size_t)H * W, o = (size_t)h * W + w;
#pragma unroll
  for (int c = 0; c < 9; ++c) ws[c * plane + o] = m[c];
}

// generic horizontal / vertical box sums over 9 fp64 channels
__global__ void box_h_kernel(const double* __restrict__ in, int H, int W, int r, double* __restrict__ out) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int h = blockIdx.y;
  if (w >= W) return;
  size_t plane = (size_t)H * W;
  int lo = max(0, w - r), hi = min(W - 1, w + r);
  for (int c = 0; c < 9; ++c) {
    const double* p = in + c * plane + (size_t)h * W;
    double s = 0;
    for (int j = lo; j <= hi; ++j) s += p[j];
    out[c * plane + (size_t)h * W + w] = s;
  }
}

__device__ __forceinline__ void box_v(const double* __restrict__ in, int H, int W, int r, int h, int w, double (&m)[9]) {
  size_t plane = (size_t)H * W;
  int lo = max(0, h - r), hi = min(H - 1, h + r);
#pragma unroll
  for (int c = 0; c < 9; ++c) m[c] = 0;
  for (int i = lo; i <= hi; ++i) {
    size_t o = (size_t)i * W + w;
#pragma unroll
    for (int c = 0; c < 9; ++c) m[c] += in[c * plane + o];
  }
}

// solves the symmetric 3x3 system M v = b by the adjugate (what torch.linalg.inv amounts to, in fp64)
__device__ __forceinline__ void solve_sym3(const double (&m)[9], double bx, double by, double bz, double& vx, double& vy,
                                           double& vz) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
  double A00 = d * f - e * e, A01 = c * e - b * f, A02 = b * e - c * d;
  double A11 = a * f - c * c, A12 = b * c - a * e, A22 = a * d - b * b;
  double det = a * A00 + b * A01 + c * A02;
  double inv = 1.0 / det;
  vx = (A00 * bx + A01 * by + A02 * bz) * inv;
  vy = (A01 * bx + A11 * by + A12 * bz) * inv;
  vz = (A02 * bx + A12 * by + A22 * bz) * inv;
}

// pass B (forward): vertical sums + solve -> normal [3][H][W]
__global__ void normal_solve_kernel(const double* __restrict__ ws, int H, int W, int r, float* __restrict__ normal) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int h = blockIdx.y;
  if (w >= W) return;
  double m[9];
  box_v(ws, H, W, r, h, w, m);
  double nx, ny, nz;
  solve_sym3(m, m[6], m[7], m[8], nx, ny, nz);
  size_t plane = (size_t)H * W, o = (size_t)h * W + w;
  normal[o] = (float)nx;
  normal[plane + o] = (float)ny;
  normal[2 * plane + o] = (float)nz;
}

// pass B (backward): vertical sums + solve n and q = M^-1 g -> 9 channels (q, sym(q n^T))
__global__ void normal_adjoint_kernel(const double* __restrict__ ws, const float* __restrict__ g_normal, int H, int W,
                                      int r, double* __restrict__ out) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int h = blockIdx.y;
  if (w >= W) return;
  double m[9];
  box_v(ws, H, W, r, h, w, m);
  double nx, ny, nz, qx, qy, qz;
  solve_sym3(m, m[6], m[7], m[8], nx, ny, nz);
  size_t plane = (size_t)H * W, o = (size_t)h * W + w;
  solve_sym3(m, (double)__ldg(g_normal + o), (double)__ldg(g_normal + plane + o), (double)__ldg(g_normal + 2 * plane + o),
             qx, qy, qz);
  out[0 * plane + o] = 2.0 * qx * nx;       // S_xx
  out[1 * plane + o] = qx * ny + nx * qy;   // S_xy
  out[2 * plane + o] = qx * nz + nx * qz;   // S_xz
  out[3 * plane + o] = 2.0 * qy * ny;       // S_yy
  out[4 * plane + o] = qy * nz + ny * qz;   // S_yz
  out[5 * plane + o] = 2.0 * qz * nz;       // S_zz
  out[6 * plane + o] = qx;
  out[7 * plane + o] = qy;
  out[8 * plane + o] = qz;
}

// pass D (backward): vertical sums of the adjoint channels -> d depth
__global__ void normal_ddepth_kernel(const double* __restrict__ ws, const float* __restrict__ depth, int H, int W, Cam cam,
                                     int r, float* __restrict__ d_depth) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int h = blockIdx.y;
  if (w >= W) return;
  double m[9];
  box_v(ws, H, W, r, h, w, m);
  float xf, yf, zf;
  xyz_of(depth, W, h, w, cam, xf, yf, zf);
  double x = xf, y = yf, z = zf;
  double dax = m[6] - (m[0] * x + m[1] * y + m[2] * z);
  double day = m[7] - (m[1] * x + m[3] * y + m[4] * z);
  double daz = m[8] - (m[2] * x + m[4] * y + m[5] * z);
  if (cam.xyz) {  // gradient w.r.t. the xyz planes themselves
    const size_t plane = (size_t)H * W, o = (size_t)h * W + w;
    d_depth[o] = (float)dax; d_depth[plane + o] = (float)day; d_depth[2 * plane + o] = (float)daz;
    return;
  }
  double g = dax * ((double)w - cam.cx) / cam.fx + day * ((double)h - cam.cy) / cam.fy + daz;
  d_depth[(size_t)h * W + w] = (float)g;
}

int check_args(const char* who, const float* depth, int H, int W, int k, const void* out, const void* workspace) {
  MVIP_REQUIRE(depth && out && workspace, MVIP_E_INVALID, "%s: null pointer", who);
  MVIP_REQUIRE(H >= 1 && W >= 1 && H <= 65535, MVIP_E_INVALID, "%s: bad image size %dx%d", who, H, W);
  MVIP_REQUIRE(k >= 1 && (k & 1) == 1, MVIP_E_UNSUPPORTED, "%s: window k=%d must be odd", who, k);
  MVIP_REQUIRE(mvip_aligned(workspace, 8), MVIP_E_INVALID, "%s: workspace must be 8-byte aligned", who);
  return MVIP_OK;
}

}  // namespace

extern "C" {

size_t mvip_normal_workspace_bytes(int H, int W) { return (size_t)2 * 9 * (size_t)H * (size_t)W * sizeof(double); }

int mvip_normal_forward_xyz(const float* xyz, int H, int W, int k, float* normal, void* workspace, void* stream) {
  int rc = check_args("mvip_normal_forward_xyz", xyz, H, W, k, normal, workspace);
  if (rc) return rc;
  Cam cam{1.f, 1.f, 0.f, 0.f, xyz, (size_t)H * W};
  dim3 block(128), grid((W + 127) / 128, H);
  double* ws = static_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  moments_h_kernel<<<grid, block, 0, st>>>(xyz, H, W, cam, k / 2, ws);
  normal_solve_kernel<<<grid, block, 0, st>>>(ws, H, W, k / 2, normal);
  MVIP_LAUNCH_OK("normal_forward_xyz kernels");
  return MVIP_OK;
}

int mvip_normal_backward_xyz(const float* xyz, int H, int W, int k, const float* g_normal, float* d_xyz, void* workspace,
                             void* stream) {
  int rc = check_args("mvip_normal_backward_xyz", xyz, H, W, k, d_xyz, workspace);
  if (rc) return rc;
  MVIP_REQUIRE(g_normal, MVIP_E_INVALID, "mvip_normal_backward_xyz: null g_normal");
  Cam cam{1.f, 1.f, 0.f, 0.f, xyz, (size_t)H * W};
  dim3 block(128), grid((W + 127) / 128, H);
  double* ws0 = static_cast<double*>(workspace);
  double* ws1 = ws0 + (size_t)9 * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  moments_h_kernel<<<grid, block, 0, st>>>(xyz, H, W, cam, k / 2, ws0);
  normal_adjoint_kernel<<<grid, block, 0, st>>>(ws0, g_normal, H, W, k / 2, ws1);
  box_h_kernel<<<grid, block, 0, st>>>(ws1, H, W, k / 2, ws0);
  normal_ddepth_kernel<<<grid, block, 0, st>>>(ws0, xyz, H, W, cam, k / 2, d_xyz);
  MVIP_LAUNCH_OK("normal_backward_xyz kernels");
  return MVIP_OK;
}

int mvip_normal_forward(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int k, float* normal,
                        void* workspace, void* stream) {
  int rc = check_args("mvip_normal_forward", depth, H, W, k, normal, workspace);
  if (rc) return rc;
  Cam cam{fx, fy, cx, cy, nullptr, 0};
  dim3 block(128), grid((W + 127) / 128, H);
  double* ws = static_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  moments_h_kernel<<<grid, block, 0, st>>>(depth, H, W, cam, k / 2, ws);
  normal_solve_kernel<<<grid, block, 0, st>>>(ws, H, W, k / 2, normal);
  MVIP_LAUNCH_OK("normal_forward kernels");
  return MVIP_OK;
}

int mvip_normal_backward(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int k,
                         const float* g_normal, float* d_depth, void* workspace, void* stream) {
  int rc = check_args("mvip_normal_backward", depth, H, W, k, d_depth, workspace);
  if (rc) return rc;
  MVIP_REQUIRE(g_normal, MVIP_E_INVALID, "mvip_normal_backward: null g_normal");
  Cam cam{fx, fy, cx, cy, nullptr, 0};
  dim3 block(128), grid((W + 127) / 128, H);
  double* ws0 = static_cast<double*>(workspace);
  double* ws1 = ws0 + (size_t)9 * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  moments_h_kernel<<<grid, block, 0, st>>>(depth, H, W, cam, k / 2, ws0);
  normal_adjoint_kernel<<<grid, block, 0, st>>>(ws0, g_normal, H, W, k / 2, ws1);
  box_h_kernel<<<grid, block, 0, st>>>(ws1, H, W, k / 2, ws0);
  normal_ddepth_kernel<<<grid, block, 0, st>>>(ws0, depth, H, W, cam, k / 2, d_depth);
  MVIP_LAUNCH_OK("normal_backward kernels");
  return MVIP_OK;
}

}  // extern "C"
