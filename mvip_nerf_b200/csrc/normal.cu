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

constexpr int kNT = 128;     // threads per block = image columns per block
constexpr int kStrip = 8;    // rows per thread in the vertical passes

__device__ __forceinline__ void moments_of(float xf, float yf, float zf, double (&m)[9]) {
  const double x = xf, y = yf, z = zf;   // products of two fp32 values are exact in fp64
  m[0] = x * x; m[1] = x * y; m[2] = x * z; m[3] = y * y; m[4] = y * z; m[5] = z * z;
  m[6] = x; m[7] = y; m[8] = z;
}

// pass A: horizontal window sums of the 9 moment channels of a -> ws[9][H][W] (fp64).
// The block's row segment (+ r halo columns each side, zero outside the image) is evaluated ONCE per pixel into shared
// memory — back-projection with its two IEEE divisions included — and every thread then adds its 2r+1 window entries in
// ascending column order (the first version re-evaluated 31 pixels per output pixel).
__global__ void __launch_bounds__(kNT) moments_h_kernel(const float* __restrict__ depth, int H, int W, Cam cam, int r,
                                                        double* __restrict__ ws) {
  extern __shared__ double tile[];   // [9][span]
  const int span = kNT + 2 * r, w0 = blockIdx.x * kNT, h = blockIdx.y;
  for (int j = threadIdx.x; j < span; j += kNT) {
    const int w = w0 - r + j;
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (w >= 0 && w < W) {
      float xf, yf, zf;
      xyz_of(depth, W, h, w, cam, xf, yf, zf);
      moments_of(xf, yf, zf, m);
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) tile[c * span + j] = m[c];
  }
  __syncthreads();
  const int w = w0 + threadIdx.x;
  if (w >= W) return;
  const size_t plane = (size_t)H * W, o = (size_t)h * W + w;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const double* t = tile + c * span + threadIdx.x;
    double s = 0;
    for (int j = 0; j <= 2 * r; ++j) s += t[j];
    ws[c * plane + o] = s;
  }
}

// generic horizontal box sums over 9 fp64 channels, staged the same way
__global__ void __launch_bounds__(kNT) box_h_kernel(const double* __restrict__ in, int H, int W, int r, double* __restrict__ out) {
  extern __shared__ double tile[];   // [9][span]
  const int span = kNT + 2 * r, w0 = blockIdx.x * kNT, h = blockIdx.y;
  const size_t plane = (size_t)H * W;
  for (int j = threadIdx.x; j < span; j += kNT) {
    const int w = w0 - r + j;
    const bool in_img = w >= 0 && w < W;
#pragma unroll
    for (int c = 0; c < 9; ++c) tile[c * span + j] = in_img ? in[c * plane + (size_t)h * W + w] : 0.0;
  }
  __syncthreads();
  const int w = w0 + threadIdx.x;
  if (w >= W) return;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const double* t = tile + c * span + threadIdx.x;
    double s = 0;
    for (int j = 0; j <= 2 * r; ++j) s += t[j];
    out[c * plane + (size_t)h * W + w] = s;
  }
}

// Vertical window sums as a SLIDING window: a thread owns kStrip consecutive rows of one column, builds the window of the first
// row once and then adds the entering / subtracts the leaving row (fp64: the drift over a strip is ~1e-16 relative, far below
// the conditioning of M) — (2r + 2 kStrip) / kStrip loads per pixel and channel instead of 2r + 1, which was L2-bandwidth bound.
// op(h, w, m) consumes the 9 window sums of pixel (h, w).
template <class Op>
__global__ void __launch_bounds__(kNT) strip_v_kernel(const double* __restrict__ in, int H, int W, int r, Op op) {
  const int w = blockIdx.x * kNT + threadIdx.x;
  const int h0 = blockIdx.y * kStrip;
  if (w >= W) return;
  const size_t plane = (size_t)H * W;
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = max(0, h0 - r); i <= min(H - 1, h0 + r); ++i) {
    const size_t o = (size_t)i * W + w;
#pragma unroll
    for (int c = 0; c < 9; ++c) m[c] += in[c * plane + o];
  }
  const int h1 = min(H, h0 + kStrip);
  for (int h = h0; h < h1; ++h) {
    op(h, w, m);
    if (h + 1 < h1) {
      const int add = h + 1 + r, sub = h - r;
      if (add < H) {
        const size_t o = (size_t)add * W + w;
#pragma unroll
        for (int c = 0; c < 9; ++c) m[c] += in[c * plane + o];
      }
      if (sub >= 0) {
        const size_t o = (size_t)sub * W + w;
#pragma unroll
        for (int c = 0; c < 9; ++c) m[c] -= in[c * plane + o];
      }
    }
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

// pass B (forward): solve -> normal [3][H][W]
struct SolveOp {
  float* normal;
  int H, W;
  __device__ void operator()(int h, int w, const double (&m)[9]) const {
    double nx, ny, nz;
    solve_sym3(m, m[6], m[7], m[8], nx, ny, nz);
    const size_t plane = (size_t)H * W, o = (size_t)h * W + w;
    normal[o] = (float)nx;
    normal[plane + o] = (float)ny;
    normal[2 * plane + o] = (float)nz;
  }
};

// pass B (backward): solve n and q = M^-1 g -> 9 channels (sym(q n^T), q)
struct AdjointOp {
  const float* g_normal;
  double* out;
  int H, W;
  __device__ void operator()(int h, int w, const double (&m)[9]) const {
    double nx, ny, nz, qx, qy, qz;
    solve_sym3(m, m[6], m[7], m[8], nx, ny, nz);
    const size_t plane = (size_t)H * W, o = (size_t)h * W + w;
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
};

// pass D (backward): window sums of the adjoint channels -> d depth (or d xyz)
struct DDepthOp {
  const float* depth;
  float* d_depth;
  Cam cam;
  int H, W;
  __device__ void operator()(int h, int w, const double (&m)[9]) const {
    float xf, yf, zf;
    xyz_of(depth, W, h, w, cam, xf, yf, zf);
    const double x = xf, y = yf, z = zf;
    const double dax = m[6] - (m[0] * x + m[1] * y + m[2] * z);
    const double day = m[7] - (m[1] * x + m[3] * y + m[4] * z);
    const double daz = m[8] - (m[2] * x + m[4] * y + m[5] * z);
    if (cam.xyz) {  // gradient w.r.t. the xyz planes themselves
      const size_t plane = (size_t)H * W, o = (size_t)h * W + w;
      d_depth[o] = (float)dax; d_depth[plane + o] = (float)day; d_depth[2 * plane + o] = (float)daz;
      return;
    }
    const double g = dax * ((double)w - cam.cx) / cam.fx + day * ((double)h - cam.cy) / cam.fy + daz;
    d_depth[(size_t)h * W + w] = (float)g;
  }
};

// launches: forward = moments_h + strip_v<SolveOp>; backward = moments_h + strip_v<AdjointOp> + box_h + strip_v<DDepthOp>
int run_forward(const float* src, int H, int W, const Cam& cam, int k, float* normal, double* ws, cudaStream_t st) {
  const int r = k / 2;
  const size_t smem = (size_t)9 * (kNT + 2 * r) * sizeof(double);
  MVIP_REQUIRE(smem <= 48 * 1024, MVIP_E_UNSUPPORTED, "normal map: window k=%d too large (k <= 553)", k);
  const dim3 block(kNT), grid_h((W + kNT - 1) / kNT, H), grid_v((W + kNT - 1) / kNT, (H + kStrip - 1) / kStrip);
  moments_h_kernel<<<grid_h, block, smem, st>>>(src, H, W, cam, r, ws);
  strip_v_kernel<SolveOp><<<grid_v, block, 0, st>>>(ws, H, W, r, SolveOp{normal, H, W});
  MVIP_LAUNCH_OK("normal forward kernels");
  return MVIP_OK;
}

int run_backward(const float* src, int H, int W, const Cam& cam, int k, const float* g_normal, float* d_out, double* ws0,
                 cudaStream_t st) {
  const int r = k / 2;
  const size_t smem = (size_t)9 * (kNT + 2 * r) * sizeof(double);
  MVIP_REQUIRE(smem <= 48 * 1024, MVIP_E_UNSUPPORTED, "normal map: window k=%d too large (k <= 553)", k);
  double* ws1 = ws0 + (size_t)9 * H * W;
  const dim3 block(kNT), grid_h((W + kNT - 1) / kNT, H), grid_v((W + kNT - 1) / kNT, (H + kStrip - 1) / kStrip);
  moments_h_kernel<<<grid_h, block, smem, st>>>(src, H, W, cam, r, ws0);
  strip_v_kernel<AdjointOp><<<grid_v, block, 0, st>>>(ws0, H, W, r, AdjointOp{g_normal, ws1, H, W});
  box_h_kernel<<<grid_h, block, smem, st>>>(ws1, H, W, r, ws0);
  strip_v_kernel<DDepthOp><<<grid_v, block, 0, st>>>(ws0, H, W, r, DDepthOp{src, d_out, cam, H, W});
  MVIP_LAUNCH_OK("normal backward kernels");
  return MVIP_OK;
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
  const Cam cam{1.f, 1.f, 0.f, 0.f, xyz, (size_t)H * W};
  return run_forward(xyz, H, W, cam, k, normal, static_cast<double*>(workspace), (cudaStream_t)stream);
}

int mvip_normal_backward_xyz(const float* xyz, int H, int W, int k, const float* g_normal, float* d_xyz, void* workspace,
                             void* stream) {
  int rc = check_args("mvip_normal_backward_xyz", xyz, H, W, k, d_xyz, workspace);
  if (rc) return rc;
  MVIP_REQUIRE(g_normal, MVIP_E_INVALID, "mvip_normal_backward_xyz: null g_normal");
  const Cam cam{1.f, 1.f, 0.f, 0.f, xyz, (size_t)H * W};
  return run_backward(xyz, H, W, cam, k, g_normal, d_xyz, static_cast<double*>(workspace), (cudaStream_t)stream);
}

int mvip_normal_forward(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int k, float* normal,
                        void* workspace, void* stream) {
  int rc = check_args("mvip_normal_forward", depth, H, W, k, normal, workspace);
  if (rc) return rc;
  const Cam cam{fx, fy, cx, cy, nullptr, 0};
  return run_forward(depth, H, W, cam, k, normal, static_cast<double*>(workspace), (cudaStream_t)stream);
}

int mvip_normal_backward(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int k,
                         const float* g_normal, float* d_depth, void* workspace, void* stream) {
  int rc = check_args("mvip_normal_backward", depth, H, W, k, d_depth, workspace);
  if (rc) return rc;
  MVIP_REQUIRE(g_normal, MVIP_E_INVALID, "mvip_normal_backward: null g_normal");
  const Cam cam{fx, fy, cx, cy, nullptr, 0};
  return run_backward(depth, H, W, cam, k, g_normal, d_depth, static_cast<double*>(workspace), (cudaStream_t)stream);
}

}  // extern "C"
