/*
 * nerf_oracle_c.c — TEST INFRASTRUCTURE ONLY: plain-C restatement of the integer / bit-exact stages of the reference's
 * render path and of its fp32 compositing, independent of numpy and of torch.  A second oracle beside oracle/nerf_oracle.py:
 * tests/test_oracle_c.py holds both against the golden vectors the reference produced (tests/golden/, oracle/make_golden*.py),
 * bit for bit where the stage is bit-exact.  Nothing under mvip_nerf_b200/ links or loads this file.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math — x86-64 SSE arithmetic on `float` IS IEEE binary32 with one
 * rounding per operation, which is what the reference's fp32 torch ops do; contraction into FMA must stay off.
 *
 * Reference lines (relative to /root/reference/DS_NeRF):
 *   mvo_ray_batch        run_nerf_helpers.py:249-260 (get_rays), :283-300 (ndc_rays), run.py:1171-1207 (batch assembly)
 *   mvo_sample_coarse    run.py:1759-1781
 *   mvo_sample_fine      run.py:1809-1816 + run_nerf_helpers.py:304-347 (sample_pdf)
 *   mvo_raw2outputs      run_nerf_helpers.py:350-404
 * Third-party rounding reproduced (PyTorch CPU): torch.sum lane order, fp64 cumsum / cumprod, upper-bound search.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* torch.sum over K contiguous floats, ATen CPU order (8 lanes x 4 accumulators); valid for K in [1,4] U [8,512] */
static float aten_row_sum(const float* x, int K) {
  const int nv = K / 8, nfull = 4 * (nv / 4);
  float ps[4][8];
  int started[4] = {0, 0, 0, 0};
  memset(ps, 0, sizeof ps);
  for (int v = 0; v < nfull; ++v) {
    const int u = v % 4;
    for (int l = 0; l < 8; ++l) ps[u][l] = started[u] ? ps[u][l] + x[8 * v + l] : x[8 * v + l];
    started[u] = 1;
  }
  for (int v = nfull; v < nv; ++v) {
    for (int l = 0; l < 8; ++l) ps[0][l] = started[0] ? ps[0][l] + x[8 * v + l] : x[8 * v + l];
    started[0] = 1;
  }
  float t[8];
  for (int l = 0; l < 8; ++l) {
    t[l] = ps[0][l];
    if (nfull > 0) { t[l] = t[l] + ps[1][l]; t[l] = t[l] + ps[2][l]; t[l] = t[l] + ps[3][l]; }
  }
  float acc = 0.f;
  for (int k = 8 * nv; k < K; ++k) acc = acc + x[k];
  if (nv > 0)
    for (int l = 0; l < 8; ++l) acc = acc + t[l];
  return acc;
}

static int upper_bound_f(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static int cmp_float(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

/* pinhole rays of the window [i0,i0+h) x [j0,j0+w) of an H x W view -> out[h*w][8 + 3*use_viewdirs] */
void mvo_ray_batch(const float* c2w, const float* c2w_static, int H, int W, double focal_d, float near, float far, int i0,
                   int j0, int h, int w, int use_viewdirs, int ndc, double ndc_near, float* out) {
  const float focal = (float)focal_d;
  const float* ps = c2w_static ? c2w_static : c2w;
  const int stride = use_viewdirs ? 11 : 8;
  const float sx = (float)(-1. / ((double)W / (2. * focal_d))), sy = (float)(-1. / ((double)H / (2. * focal_d)));
  const float nn = (float)ndc_near, two_n = (float)(2. * ndc_near), neg_two_n = (float)(-2. * ndc_near);
  for (int a = 0; a < h; ++a)
    for (int b = 0; b < w; ++b) {
      const int i = i0 + a, j = j0 + b;
      const float dx = ((float)j - (float)W * .5f) / focal;
      const float dy = -(((float)i - (float)H * .5f) / focal);
      const float dz = -1.f;
      float ro[3], rd[3], v[3];
      for (int k = 0; k < 3; ++k) {
        ro[k] = ps[4 * k + 3];
        rd[k] = (dx * ps[4 * k] + dy * ps[4 * k + 1]) + dz * ps[4 * k + 2];
        v[k] = (dx * c2w[4 * k] + dy * c2w[4 * k + 1]) + dz * c2w[4 * k + 2];
      }
      float* o = out + ((size_t)a * w + b) * stride;
      if (ndc) {
        const float t = -(nn + ro[2]) / rd[2];
        const float ox = ro[0] + t * rd[0], oy = ro[1] + t * rd[1], oz = ro[2] + t * rd[2];
        const float rz = 1.f / oz;
        o[0] = (sx * ox) / oz;
        o[1] = (sy * oy) / oz;
        o[2] = 1.f + rz * two_n;
        o[3] = sx * (rd[0] / rd[2] - ox / oz);
        o[4] = sy * (rd[1] / rd[2] - oy / oz);
        o[5] = rz * neg_two_n;
      } else {
        for (int k = 0; k < 3; ++k) { o[k] = ro[k]; o[3 + k] = rd[k]; }
      }
      o[6] = near;
      o[7] = far;
      if (use_viewdirs) {   /* torch.norm: sqrt(fma(z, z, fma(y, y, x*x))) */
        const float nrm = sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0])));
        o[8] = v[0] / nrm; o[9] = v[1] / nrm; o[10] = v[2] / nrm;
      }
    }
}

void mvo_sample_coarse(const float* rays, int stride, int64_t n, const float* t_vals, const float* t_rand, int S, int lindisp,
                       float* z_out) {
  float* z = (float*)malloc(sizeof(float) * (size_t)S);
  for (int64_t r = 0; r < n; ++r) {
    const float near = rays[r * stride + 6], far = rays[r * stride + 7];
    for (int i = 0; i < S; ++i) {
      const float t = t_vals[i], omt = 1.0f - t;
      if (!lindisp) {
        z[i] = near * omt + far * t;
      } else {
        const float a = 1.0f / near, b = 1.0f / far;
        const float den = a * omt + b * t;
        z[i] = 1.0f / den;
      }
    }
    float* out = z_out + r * S;
    if (!t_rand) {
      memcpy(out, z, sizeof(float) * (size_t)S);
      continue;
    }
    for (int i = 0; i < S; ++i) {
      const float lower = i > 0 ? .5f * (z[i] + z[i - 1]) : z[0];
      const float upper = i < S - 1 ? .5f * (z[i + 1] + z[i]) : z[S - 1];
      out[i] = lower + (upper - lower) * t_rand[r * S + i];
    }
  }
  free(z);
}

/* z [n,S], weights [n,S] (raw2outputs weights of the coarse pass), u [n,M] or one row [M] ->
 * samples [n,M] (nullable), inds [n,M] int64 (nullable), cdf [n,S-1] (nullable), merged [n,S+M] */
void mvo_sample_fine(const float* z_g, const float* weights_g, const float* u_g, int u_is_row, int64_t n, int S, int M,
                     float* samples_g, int64_t* inds_g, float* cdf_g, float* merged_g) {
  const int B = S - 1, K = S - 2;
  float* bins = (float*)malloc(sizeof(float) * (size_t)(3 * S + M + S));
  float* w = bins + S;
  float* cdf = w + S;
  float* all = cdf + S;     /* S + M */
  for (int64_t r = 0; r < n; ++r) {
    const float* z = z_g + r * S;
    for (int i = 0; i < B; ++i) bins[i] = .5f * (z[i + 1] + z[i]);
    for (int i = 0; i < K; ++i) w[i] = weights_g[r * S + 1 + i] + 1e-5f;
    const float s = aten_row_sum(w, K);
    double run = 0.0;
    cdf[0] = 0.f;
    for (int i = 0; i < K; ++i) {
      const float p = w[i] / s;
      run += (double)p;
      cdf[i + 1] = (float)run;
    }
    if (cdf_g) memcpy(cdf_g + r * B, cdf, sizeof(float) * (size_t)B);
    memcpy(all, z, sizeof(float) * (size_t)S);
    for (int j = 0; j < M; ++j) {
      const float u = u_is_row ? u_g[j] : u_g[r * M + j];
      const int ind = upper_bound_f(cdf, B, u);
      const int below = ind - 1 > 0 ? ind - 1 : 0, above = ind < B - 1 ? ind : B - 1;
      const float cb = cdf[below];
      float denom = cdf[above] - cb;
      if (denom < 1e-5f) denom = 1.0f;
      const float t = (u - cb) / denom;
      const float bb = bins[below];
      const float smp = bb + t * (bins[above] - bb);
      if (samples_g) samples_g[r * M + j] = smp;
      if (inds_g) inds_g[r * M + j] = ind;
      all[S + j] = smp;
    }
    qsort(all, (size_t)(S + M), sizeof(float), cmp_float);      /* torch.sort(cat[z, z_samples]) — values only */
    memcpy(merged_g + r * (S + M), all, sizeof(float) * (size_t)(S + M));
  }
  free(bins);
}

/* raw [n,S,4], z [n,S], rays_d [n,3], noise [n,S] or NULL -> rgb [n,3], disp, acc, depth [n], weights [n,S], alpha [n,S] (nullable) */
void mvo_raw2outputs(const float* raw, const float* z, const float* rays_d, const float* noise, int64_t n, int S, int white,
                     float* rgb, float* disp, float* acc, float* depth, float* weights, float* alpha) {
  for (int64_t r = 0; r < n; ++r) {
    const float* d = rays_d + r * 3;
    const float nrm = sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0])));
    double T = 1.0;     /* torch.cumprod accumulates in fp64 and rounds every output to fp32 */
    float sR = 0.f, sG = 0.f, sB = 0.f, sD = 0.f, sA = 0.f;
    for (int i = 0; i < S; ++i) {
      const float* q = raw + (r * S + i) * 4;
      const float dist = (i + 1 < S ? z[r * S + i + 1] - z[r * S + i] : 1e10f) * nrm;
      float sig = q[3] + (noise ? noise[r * S + i] : 0.f);
      if (sig < 0.f) sig = 0.f;
      const float a = 1.f - expf(-sig * dist);
      const float Ti = (float)T;          /* exclusive product: cat[1, 1-a+1e-10][:-1] */
      const float wgt = a * Ti;
      T *= (double)((1.f - a) + 1e-10f);
      weights[r * S + i] = wgt;
      if (alpha) alpha[r * S + i] = a;
      sR += wgt * (1.f / (1.f + expf(-q[0])));
      sG += wgt * (1.f / (1.f + expf(-q[1])));
      sB += wgt * (1.f / (1.f + expf(-q[2])));
      sD += wgt * z[r * S + i];
      sA += wgt;
    }
    const float ratio = sD / sA;
    const float m = (ratio != ratio) ? ratio : (ratio > 1e-10f ? ratio : 1e-10f);   /* torch.max propagates NaN */
    const float bg = white ? 1.f - sA : 0.f;
    rgb[r * 3] = sR + bg; rgb[r * 3 + 1] = sG + bg; rgb[r * 3 + 2] = sB + bg;
    disp[r] = 1.f / m;
    acc[r] = sA;
    depth[r] = sD;
  }
}
