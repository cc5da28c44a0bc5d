// sampler.cu — stratified sampling and hierarchical (inverse-CDF) sampling, bit-exact against the
// reference's CPU path.  Compiled with -fmad=false; every rounding step below is explicit.
//
//   mvip_sample_coarse : DS_NeRF/run.py:1759-1781
//   mvip_sample_pdf    : DS_NeRF/run_nerf_helpers.py:304-347
//   mvip_sample_fine   : DS_NeRF/run.py:1809-1816, 1836
//
// Third-party rounding that is reproduced here (re-verified in oracle/nerf_oracle.py):
//   torch.sum(x,-1) on CPU  = 8 SIMD lanes x 4 interleaved accumulators (aten_row_sum below)
//   torch.cumsum on CPU     = fp64 accumulation, fp32 rounding per output
//   torch.searchsorted(right=True) = upper bound, int64
#include "common.cuh"

#include <math_constants.h>

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxBins = 256;  // n_bins (= N_samples-1 inside render_rays)
constexpr int kMaxOut = 256;   // N_importance

// ---------------------------------------------------------------------------------------------
__global__ void sample_coarse_kernel(const float* __restrict__ rays, int ray_stride, int64_t n_rays,
                                     const float* __restrict__ t_vals, const float* __restrict__ t_rand, int S,
                                     int lindisp, float* __restrict__ z_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n_rays * S;
  if (idx >= total) return;
  int64_t r = idx / S;
  int i = (int)(idx - r * S);
  float near = __ldg(rays + r * ray_stride + 6);
  float far = __ldg(rays + r * ray_stride + 7);
  float inv_near = 0.f, inv_far = 0.f;
  if (lindisp) {
    inv_near = __fdiv_rn(1.0f, near);
    inv_far = __fdiv_rn(1.0f, far);
  }
  auto base = [&](int j) -> float {
    float t = __ldg(t_vals + j);
    float omt = __fsub_rn(1.0f, t);
    if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
    float den = __fadd_rn(__fmul_rn(inv_near, omt), __fmul_rn(inv_far, t));
    return __fdiv_rn(1.0f, den);
  };
  float z = base(i);
  if (t_rand != nullptr) {
    float lower = z, upper = z;
    if (i > 0) lower = __fmul_rn(0.5f, __fadd_rn(z, base(i - 1)));
    if (i < S - 1) upper = __fmul_rn(0.5f, __fadd_rn(base(i + 1), z));
    float tr = __ldg(t_rand + idx);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
  }
  z_out[idx] = z;
}

// Register form for S = 4 G (G = 8, 16, 32 lanes per ray, 32 / G rays per warp): lane `sub` of a ray's group owns the FOUR
// samples [4 sub, 4 sub + 4), so t_rand / z always move as 16-byte vectors.  The per-sample table values (t, 1 - t) live in
// registers across rays, 1/near and 1/far are ONE reciprocal instruction (lane parity picks the operand), every stratum is
// evaluated once and its neighbours' values arrive by shuffle inside the group.  ~1.5 instructions per sample instead of
// ~120 (three stratum evaluations with five IEEE divisions each).
// __frcp_rn is the correctly rounded reciprocal == IEEE 1/x == torch's reciprocal.
template <int G>
__global__ void __launch_bounds__(256)
sample_coarse_warp_kernel(const float* __restrict__ rays, int ray_stride, int64_t n_rays, const float* __restrict__ t_vals,
                          const float* __restrict__ t_rand, int lindisp, float* __restrict__ z_out) {
  constexpr int C = 4, S = C * G, RPW = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
  float t[C], omt[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    t[k] = __ldg(t_vals + sub * C + k);
    omt[k] = __fsub_rn(1.0f, t[k]);
  }
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // kU ray groups per trip: all their loads (near / far, the draws) are issued before the first use
  constexpr int kU = (G == 32) ? 4 : 1;      // S = 64 (two rays per warp), 262,144 rays: kU = 1 / 2 / 4 / 8 -> 65.9 / 58.3 / 52.3 / 47.4 % of the copy peak
  for (int64_t r0 = warp0 * RPW; r0 < n_rays; r0 += kU * nwarps * RPW) {
    float nf[kU], tr[kU][C];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t r = r0 + u * nwarps * RPW + grp;
      const bool live = r < n_rays;
      nf[u] = live ? __ldg(rays + r * ray_stride + 6 + (sub & 1)) : 1.f;     // even lanes: near, odd lanes: far
      if (t_rand != nullptr && live) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(t_rand + r * S) + sub);
        tr[u][0] = v.x; tr[u][1] = v.y; tr[u][2] = v.z; tr[u][3] = v.w;
      } else {
#pragma unroll
        for (int k = 0; k < C; ++k) tr[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t rg0 = r0 + u * nwarps * RPW;
      if (rg0 >= n_rays) break;                                    // warp-uniform
      const int64_t r = rg0 + grp;
      float nfu = nf[u];
      if (lindisp) nfu = __frcp_rn(nfu);
      const float a = __shfl_sync(FULL_MASK, nfu, 0, G), b = __shfl_sync(FULL_MASK, nfu, 1, G);
      float z[C];
#pragma unroll
      for (int k = 0; k < C; ++k) {
        z[k] = __fadd_rn(__fmul_rn(a, omt[k]), __fmul_rn(b, t[k]));
        if (lindisp) z[k] = __frcp_rn(z[k]);
      }
      if (t_rand != nullptr) {
        const float prev = __shfl_up_sync(FULL_MASK, z[C - 1], 1, G), next = __shfl_down_sync(FULL_MASK, z[0], 1, G);
        float out[C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const float zp = (k > 0) ? z[(k > 0) ? k - 1 : 0] : prev, zn = (k + 1 < C) ? z[(k + 1 < C) ? k + 1 : k] : next;
          float lower = __fmul_rn(0.5f, __fadd_rn(z[k], zp)), upper = __fmul_rn(0.5f, __fadd_rn(zn, z[k]));
          if (k == 0 && sub == 0) lower = z[k];
          if (k == C - 1 && sub == G - 1) upper = z[k];
          out[k] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr[u][k]));
        }
#pragma unroll
        for (int k = 0; k < C; ++k) z[k] = out[k];
      }
      if (r < n_rays) __stcs(reinterpret_cast<float4*>(z_out + r * S) + sub, make_float4(z[0], z[1], z[2], z[3]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// torch.sum over K contiguous floats in ATen's CPU order; every lane returns the same value.
__device__ float aten_row_sum(const float* x, int K, int lane) {
  const int nv = K >> 3;
  const int nfull = (nv >> 2) << 2;
  float t = 0.f;
  if (lane < 8) {
    float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
    for (int v = 0; v < nfull; v += 4) {
      float a0 = x[8 * v + lane], a1 = x[8 * (v + 1) + lane], a2 = x[8 * (v + 2) + lane], a3 = x[8 * (v + 3) + lane];
      if (v == 0) {
        ps0 = a0; ps1 = a1; ps2 = a2; ps3 = a3;
      } else {
        ps0 = __fadd_rn(ps0, a0); ps1 = __fadd_rn(ps1, a1); ps2 = __fadd_rn(ps2, a2); ps3 = __fadd_rn(ps3, a3);
      }
    }
    for (int v = nfull; v < nv; ++v) {
      float a = x[8 * v + lane];
      ps0 = (v == 0) ? a : __fadd_rn(ps0, a);
    }
    t = ps0;
    if (nfull > 0) t = __fadd_rn(__fadd_rn(__fadd_rn(ps0, ps1), ps2), ps3);
  }
  float acc = 0.f;
  for (int k = 8 * nv; k < K; ++k) acc = __fadd_rn(acc, x[k]);
  if (nv > 0) {
#pragma unroll
    for (int l = 0; l < 8; ++l) acc = __fadd_rn(acc, __shfl_sync(FULL_MASK, t, l));
  }
  return acc;
}

// One warp turns (bins[B], w[K=B-1] (already +1e-5)) into cdf[B] in shared memory.
// The fp64 warp scan is exact (hence order-independent, == the sequential CPU scan) whenever every
// pdf value is > 0 and >= 2^-27; otherwise lane 0 falls back to the sequential scan.
__device__ void warp_build_cdf(float* w, float* cdf, int K, int lane) {
  float s = aten_row_sum(w, K, lane);
  bool ok = true;
  for (int i = lane; i < K; i += 32) {
    float p = __fdiv_rn(w[i], s);
    w[i] = p;  // w now holds the pdf
    ok = ok && (p >= 7.450580596923828e-09f) && (p <= 1.0f);
  }
  ok = __all_sync(FULL_MASK, ok);
  __syncwarp();
  if (ok) {
    const int c = (K + 31) >> 5;
    const int b = lane * c;
    const int e = min(K, b + c);
    double local = 0.0;
    for (int i = b; i < e; ++i) local += (double)w[i];
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double v = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    double run = incl - local;
    for (int i = b; i < e; ++i) {
      run += (double)w[i];
      cdf[i + 1] = (float)run;
    }
    if (lane == 0) cdf[0] = 0.f;
  } else if (lane == 0) {
    double run = 0.0;
    cdf[0] = 0.f;
    for (int i = 0; i < K; ++i) {
      run += (double)w[i];
      cdf[i + 1] = (float)run;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u, int* ind_out) {
  int ind = upper_bound(cdf, B, u);
  *ind_out = ind;
  int below = max(0, ind - 1);
  int above = min(B - 1, ind);
  float cb = cdf[below];
  float denom = __fsub_rn(cdf[above], cb);
  if (denom < 1e-5f) denom = 1.0f;
  float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  float bb = bins[below];
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(bins[above], bb)));
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_kernel(const float* __restrict__ bins_g, const float* __restrict__ weights_g, const float* __restrict__ u_g,
                  int u_is_row, int64_t n_rows, int B, int M, float* __restrict__ samples_g,
                  int64_t* __restrict__ inds_g, float* __restrict__ cdf_g) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = B - 1;
  float* bins = smem + warp * (3 * B);
  float* w = bins + B;
  float* cdf = w + B;
  for (int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + warp; row < n_rows;
       row += (int64_t)gridDim.x * kWarpsPerBlock) {
    for (int i = lane; i < B; i += 32) bins[i] = __ldg(bins_g + row * B + i);
    for (int i = lane; i < K; i += 32) w[i] = __fadd_rn(__ldg(weights_g + row * K + i), 1e-5f);
    __syncwarp();
    warp_build_cdf(w, cdf, K, lane);
    if (cdf_g) for (int i = lane; i < B; i += 32) cdf_g[row * B + i] = cdf[i];
    for (int j = lane; j < M; j += 32) {
      float u = u_is_row ? __ldg(u_g + j) : __ldg(u_g + row * M + j);
      int ind;
      float smp = invert_cdf(cdf, bins, B, u, &ind);
      samples_g[row * M + j] = smp;
      if (inds_g) inds_g[row * M + j] = (int64_t)ind;
    }
    __syncwarp();
  }
}

// bitonic sort of P2 (power of two) floats in shared memory by one warp, ascending
__device__ void warp_bitonic_sort(float* a, int P2, int lane) {
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P2; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          float x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_fine_kernel(const float* __restrict__ z_g, const float* __restrict__ weights_g, const float* __restrict__ u_g,
                   int u_is_row, int64_t n_rays, int S, int M, int P2, float* __restrict__ samples_g,
                   int64_t* __restrict__ inds_g, float* __restrict__ merged_g, float* __restrict__ std_g) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = S - 1, K = S - 2, T = S + M;
  const int per_warp = S + 3 * B + P2 + T;
  float* z = smem + warp * per_warp;
  float* bins = z + S;
  float* w = bins + B;
  float* cdf = w + B;
  float* smp = cdf + B;   // P2 entries
  float* out = smp + P2;  // T entries
  for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp; ray < n_rays;
       ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    for (int i = lane; i < S; i += 32) z[i] = __ldg(z_g + ray * S + i);
    for (int i = lane; i < K; i += 32) w[i] = __fadd_rn(__ldg(weights_g + ray * S + 1 + i), 1e-5f);
    __syncwarp();
    for (int i = lane; i < B; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(z[i + 1], z[i]));
    __syncwarp();
    warp_build_cdf(w, cdf, K, lane);
    double sum = 0.0;
    for (int j = lane; j < P2; j += 32) {
      float v = CUDART_INF_F;
      if (j < M) {
        float u = u_is_row ? __ldg(u_g + j) : __ldg(u_g + ray * M + j);
        int ind;
        v = invert_cdf(cdf, bins, B, u, &ind);
        if (samples_g) samples_g[ray * M + j] = v;
        if (inds_g) inds_g[ray * M + j] = (int64_t)ind;
        sum += (double)v;
      }
      smp[j] = v;
    }
    __syncwarp();
    if (std_g) {  // torch.std(unbiased=False), run.py:1836 (tolerance-checked, not bit-exact)
      double mean = warp_sum(sum) / (double)M;
      double sq = 0.0;
      for (int j = lane; j < M; j += 32) {
        double d = (double)smp[j] - mean;
        sq += d * d;
      }
      sq = warp_sum(sq);
      if (lane == 0) std_g[ray] = (float)sqrt(sq / (double)M);
    }
    bool sorted = true;
    for (int j = lane; j + 1 < M; j += 32) sorted = sorted && (smp[j] <= smp[j + 1]);
    if (!__all_sync(FULL_MASK, sorted)) warp_bitonic_sort(smp, P2, lane);
    // merge two sorted runs by rank: z elements go before equal samples
    for (int i = lane; i < S; i += 32) {
      float v = z[i];
      out[i + lower_bound(smp, M, v)] = v;
    }
    for (int j = lane; j < M; j += 32) {
      float v = smp[j];
      out[j + upper_bound(z, S, v)] = v;
    }
    __syncwarp();
    for (int i = lane; i < T; i += 32) merged_g[ray * T + i] = out[i];
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// sample_fine for the shape render_rays uses (N_samples = 64 coarse depths, N_importance = 64): FOUR rays per warp, eight
// lanes per ray, lane t of a group owns the eight consecutive elements [8t, 8t+8) of every row (z, weights, draws), moved
// as 16-byte vectors.  Compared with the generic kernel:
//   * the weights go through shared memory once, for the ATen-order row sum (a transposed access); the fp64 cdf scan works
//     on register values (three shuffle steps inside the group); only cdf / bins (the search tables) live in shared memory,
//     72 floats apart per ray so that the four groups of a warp hit different banks;
//   * the searches are the branch-free 6-step form; the first three levels (cdf[31], cdf[15|47], cdf[7|23|39|55]) are
//     held in registers and shared by the lane's eight draws, the last three and the interpolation operands are loads;
//   * unsorted draws are ordered by a REGISTER bitonic network: 21 stages, 15 of them inside the lane (plain min / max on
//     registers), 6 across lanes (8 shuffles each, for four rays at once); the sorted run is merged with the coarse depths
//     by a 7-stage bitonic MERGE (reverse the samples, compare-exchange at distances 64..1: three of them across lanes).
//   A ray costs ~31 shuffle and ~25 shared-memory instructions (the two-elements-per-lane version: ~85 and ~40; the
//   shuffle / shared-memory pipe is what bounds this stage, profiles/r02_hbm_stages.md).
// Every value written is bit-identical to the generic kernel's (same roundings; sort / merge only permute values).
// ---------------------------------------------------------------------------------------------
constexpr int kF8Stride = 72;   // floats between the W1 / bins rows of neighbouring rays (64 + 8: a bank shift of 8 per group)
constexpr int kCdfStride = 65;  // ... and between their cdf rows

__device__ __forceinline__ void cmpx(float& lo, float& hi) {
  const float a = fminf(lo, hi), b = fmaxf(lo, hi);
  lo = a; hi = b;
}
// shared-memory load at a 32-bit shared address plus an immediate byte offset (ordered like a volatile access)
template <int OFF>
__device__ __forceinline__ float lds_off(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

#ifndef MVIP_SF_MINB
#define MVIP_SF_MINB 4
#endif
template <bool U_ROW>
__global__ void __launch_bounds__(256, MVIP_SF_MINB)
sample_fine64_kernel(const float* __restrict__ z_g, const float* __restrict__ weights_g, const float* __restrict__ u_g,
                     int64_t n_rays, float* __restrict__ samples_g, int64_t* __restrict__ inds_g,
                     float* __restrict__ merged_g, float* __restrict__ std_g) {
  // per warp: W1[j] = weights[j] + 1e-5 (w[i] = W1[i+1]; later the pdf for the sequential fallback), bins, and the cdf.
  // W1 / bins rows are 72 floats apart (16-byte vector stores; the transposed reads of the row sum are conflict-free);
  // the cdf rows - the table the searches probe - are 65 apart: the probes of a search level sit at equal positions
  // modulo 8 / 4 / 2, so a row offset that is a multiple of 8 banks puts the four rays of a warp on the same few
  // banks (simulated: 31.8 shared-memory cycles per draw at 72, 21.7 at 65; one dimension of scalar stores pays it).
  __shared__ __align__(16) float sm[8][3][4 * kF8Stride + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & 7, q = lane >> 3;
  float* W1 = sm[warp][0] + kF8Stride * q;
  float* CT = sm[warp][1] + kCdfStride * q;        // cdf[i] = CT[i], i = 0..62
  float* BN = sm[warp][2] + kF8Stride * q;
  float u[8];
  if (U_ROW) ldg8(u_g + 8 * t, u);
  for (int64_t ray0 = ((int64_t)blockIdx.x * 8 + warp) * 4; ray0 < n_rays; ray0 += (int64_t)gridDim.x * 32) {
    const bool active = ray0 + q < n_rays;
    const int64_t ray = active ? ray0 + q : n_rays - 1;     // idle groups shadow the last ray (nothing is stored)
    float z[8], wv[8];
    ldg8(z_g + ray * 64 + 8 * t, z);
    ldg8(weights_g + ray * 64 + 8 * t, wv);
    if (!U_ROW) ldg8(u_g + ray * 64 + 8 * t, u);
#pragma unroll
    for (int k = 0; k < 8; ++k) wv[k] = __fadd_rn(wv[k], 1e-5f);
    st8(W1 + 8 * t, wv);
    {
      const float z_next = __shfl_down_sync(FULL_MASK, z[0], 1, 8);
      float b[8];
#pragma unroll
      for (int k = 0; k < 7; ++k) b[k] = __fmul_rn(0.5f, __fadd_rn(z[k + 1], z[k]));
      b[7] = __fmul_rn(0.5f, __fadd_rn(z_next, z[7]));      // bins[63] (t = 7) does not exist and is never read
      st8(BN + 8 * t, b);
    }
    __syncwarp();
    // torch.sum in ATen's CPU order for K = 62: 7 vectors of 8 lanes (4 accumulators, vectors 4..6 fold into the first),
    // then the 6-element tail, then the 8 lanes of the combined vector, all sequential fp32 adds
    float s;
    {
      float ps0 = W1[1 + t];
      const float ps1 = W1[9 + t], ps2 = W1[17 + t], ps3 = W1[25 + t];
      ps0 = __fadd_rn(ps0, W1[33 + t]);
      ps0 = __fadd_rn(ps0, W1[41 + t]);
      ps0 = __fadd_rn(ps0, W1[49 + t]);
      const float tv = __fadd_rn(__fadd_rn(__fadd_rn(ps0, ps1), ps2), ps3);
      s = 0.f;
#pragma unroll
      for (int k = 57; k < 63; ++k) s = __fadd_rn(s, W1[k]);
#pragma unroll
      for (int l = 0; l < 8; ++l) s = __fadd_rn(s, __shfl_sync(FULL_MASK, tv, l, 8));
    }
    // pdf entries of this lane: p[8t + k] = w[8t + k] / s = W1[8t + k + 1] / s, k = 0..7 (62 entries: t = 7 has six).
    // IEEE division: the eight quotients share the divisor, so the reciprocal refinement of the compiler's own div.rn
    // sequence (MUFU.RCP, two FFMAs) is done once and each quotient takes its three remaining FFMAs - the same
    // operations on the same values, hence the same bits.  That sequence is only valid away from the exponent limits
    // (the compiler guards it with FCHK); outside a generous safe range the plain __fdiv_rn is used.
    float p[8];
    {
      const float w_next = __shfl_down_sync(FULL_MASK, wv[0], 1, 8);
      float num[8];
#pragma unroll
      for (int k = 0; k < 7; ++k) num[k] = wv[k + 1];
      num[7] = w_next;
      bool safe = (s >= 1e-30f) && (s <= 1e30f);
#pragma unroll
      for (int k = 0; k < 8; ++k) safe = safe && (num[k] >= 1e-30f) && (num[k] <= 1e30f);
      if (__all_sync(FULL_MASK, safe)) {
        float r0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s));
        const float r = __fmaf_rn(r0, __fmaf_rn(r0, -s, 1.0f), r0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float q0 = __fmaf_rn(num[k], r, 0.0f);
          p[k] = __fmaf_rn(r, __fmaf_rn(q0, -s, num[k]), q0);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = __fdiv_rn(num[k], s);
      }
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (8 * t + k < 62) ok = ok && (p[k] >= 7.450580596923828e-09f) && (p[k] <= 1.0f);
    const unsigned okb = __ballot_sync(FULL_MASK, ok);
    const bool gok = ((okb >> (8 * q)) & 0xffu) == 0xffu;
    {   // exact fp64 scan == the sequential CPU cumsum (see warp_build_cdf)
      double d[8], local = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k] = (8 * t + k < 62) ? (double)p[k] : 0.0; local += d[k]; }
      double incl = local;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const double v = __shfl_up_sync(FULL_MASK, incl, o, 8);
        if (t >= o) incl += v;
      }
      double run = incl - local;
      float c[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { run += d[k]; c[k] = (float)run; }
      if (gok) {
#pragma unroll
        for (int k = 0; k < 8; ++k) CT[8 * t + k + 1] = c[k];   // (t = 7: slots 63, 64 are padding)
      }
      if (t == 0) CT[0] = 0.f;
    }
    if (okb != 0xffffffffu) {   // a pdf entry below 2^-27 somewhere in the warp: that ray's cdf by the sequential scan
      __syncwarp();             // all reads of W1 are done
      st8(W1 + 8 * t, p);       // W1[i] = p[i] now
      __syncwarp();
      if (!gok && t == 0) {
        double run = 0.0;
        for (int i = 0; i < 62; ++i) {
          run += (double)W1[i];
          CT[i + 1] = (float)run;
        }
      }
    }
    __syncwarp();
    // ---- the draws: upper_bound(cdf, u) with the first three levels in registers, then the lerp (no FMA)
    float sv[8];
    {
      // Shared-memory addresses are kept as 32-bit byte addresses and every probe is a load with an immediate offset
      // (the compiler's own index arithmetic costs 30 instructions per draw instead of 17).
      const uint32_t cdf0 = smem_u32(CT);                   // address of cdf[0]
      const uint32_t to_bins = smem_u32(BN) - cdf0;         // bytes from cdf[i] to bins[i] of this ray
      const float c31 = lds_off<31 * 4>(cdf0), c15 = lds_off<15 * 4>(cdf0), c47 = lds_off<47 * 4>(cdf0);
      uint32_t pa[8];                                       // address of cdf[pos]; the probe of a step is cdf[pos + step - 1]
#pragma unroll
      for (int k = 0; k < 8; ++k) {                         // the first two levels from registers
        const bool hi = c31 <= u[k];
        pa[k] = cdf0 + (hi ? 128u : 0u) + (((hi ? c47 : c15) <= u[k]) ? 64u : 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) pa[k] += (lds_off<7 * 4>(pa[k]) <= u[k]) ? 32u : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) pa[k] += (lds_off<3 * 4>(pa[k]) <= u[k]) ? 16u : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) pa[k] += (lds_off<1 * 4>(pa[k]) <= u[k]) ? 8u : 0u;
      float v6[8];                                          // the last probe, cdf[pos6] (pos6 even): it IS one of the two lerp operands
      bool inc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v6[k] = lds_off<0>(pa[k]);
        inc[k] = v6[k] <= u[k];
        pa[k] += inc[k] ? 4u : 0u;
      }
      float mean_acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float uu = u[k];
        const uint32_t pl = pa[k] - (pa[k] != cdf0 ? 4u : 0u);              // &cdf[below], below = max(0, pos - 1)
        const uint32_t ph = pa[k] - (pa[k] == cdf0 + 63u * 4u ? 4u : 0u);   // &cdf[above], above = min(62, pos)
        // pos = pos6 + 1: cdf[below] is the probe, cdf[above] is loaded; pos = pos6: cdf[above] is the probe, cdf[below] is loaded
        const float other = lds_off<0>(inc[k] ? ph : pl);
        const float cb = inc[k] ? v6[k] : other;
        float denom = __fsub_rn(inc[k] ? other : v6[k], cb);
        if (denom < 1e-5f) denom = 1.0f;
        const float tt = __fdiv_rn(__fsub_rn(uu, cb), denom);
        const float bb = lds_off<0>(pl + to_bins);
        sv[k] = __fadd_rn(bb, __fmul_rn(tt, __fsub_rn(lds_off<0>(ph + to_bins), bb)));
        mean_acc += sv[k];
      }
      if (active) {
        if (samples_g) st8(samples_g + ray * 64 + 8 * t, sv);
        if (inds_g) {
          longlong2* dst = reinterpret_cast<longlong2*>(inds_g + ray * 64 + 8 * t);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            dst[k] = make_longlong2((long long)((pa[2 * k] - cdf0) >> 2), (long long)((pa[2 * k + 1] - cdf0) >> 2));
        }
      }
      if (std_g) {  // torch.std(unbiased=False), run.py:1836: fp32 two-pass like the reference's (tolerance-checked, not bit-exact)
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) mean_acc += __shfl_xor_sync(FULL_MASK, mean_acc, o);
        const float mean = mean_acc * (1.0f / 64.0f);
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float e = sv[k] - mean; sq += e * e; }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sq += __shfl_xor_sync(FULL_MASK, sq, o);
        if (t == 0 && active) std_g[ray] = sqrtf(sq * (1.0f / 64.0f));
      }
    }
    // ---- the samples in DESCENDING order (element e = 8t + k holds the (63 - e)-th smallest): the second half of the bitonic
    //      sequence the merge below starts from.  Unsorted draws: a descending sort network; ascending draws: reversed by shuffle.
    float b[8];
    {
      const float s_next = __shfl_down_sync(FULL_MASK, sv[0], 1, 8);
      bool sorted = (t == 7) || (sv[7] <= s_next);
#pragma unroll
      for (int k = 0; k < 7; ++k) sorted = sorted && (sv[k] <= sv[k + 1]);
      if (__all_sync(FULL_MASK, sorted)) {
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = __shfl_sync(FULL_MASK, sv[7 - k], 7 - t, 8);
      } else {
#pragma unroll
        for (int k2 = 2; k2 <= 64; k2 <<= 1) {
          // merge of sorted runs of k2 / 2 elements: first stage e <-> e ^ (k2 - 1) (the second run taken backwards),
          // then e <-> e ^ j for j = k2 / 4 .. 1; every comparator leaves the MAXIMUM at the lower element
          if (k2 <= 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((k & (k2 >> 1)) == 0) cmpx(sv[k ^ (k2 - 1)], sv[k]);
          } else {
            const int lm = (k2 >> 3) - 1;                   // lane t <-> t ^ lm, element k <-> 7 - k
            const bool keep_min = (t & (k2 >> 4)) == 0;
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __shfl_xor_sync(FULL_MASK, sv[7 - k], lm);
#pragma unroll
            for (int k = 0; k < 8; ++k) sv[k] = keep_min ? fmaxf(sv[k], o[k]) : fminf(sv[k], o[k]);   // keep_min = lower element: it keeps the max
          }
#pragma unroll
          for (int j = k2 >> 2; j >= 8; j >>= 1) {
            const bool keep_min = (t & (j >> 3)) == 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float o = __shfl_xor_sync(FULL_MASK, sv[k], j >> 3);
              sv[k] = keep_min ? fmaxf(sv[k], o) : fminf(sv[k], o);
            }
          }
#pragma unroll
          for (int j = (k2 >> 2) < 4 ? (k2 >> 2) : 4; j >= 1; j >>= 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((k & j) == 0) cmpx(sv[k | j], sv[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = sv[k];
      }
    }
    // ---- bitonic merge of (z ascending, samples descending): elements 8t + k (a) and 64 + 8t + k (b)
    {
#pragma unroll
      for (int k = 0; k < 8; ++k) cmpx(z[k], b[k]);
#pragma unroll
      for (int m = 4; m > 0; m >>= 1) {
        const bool keep_min = (t & m) == 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float oa = __shfl_xor_sync(FULL_MASK, z[k], m), ob = __shfl_xor_sync(FULL_MASK, b[k], m);
          z[k] = keep_min ? fminf(z[k], oa) : fmaxf(z[k], oa);
          b[k] = keep_min ? fminf(b[k], ob) : fmaxf(b[k], ob);
        }
      }
#pragma unroll
      for (int j = 4; j >= 1; j >>= 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if ((k & j) == 0) { cmpx(z[k], z[k | j]); cmpx(b[k], b[k | j]); }
      }
      if (active) {
        st8(merged_g + ray * 128 + 8 * t, z);
        st8(merged_g + ray * 128 + 64 + 8 * t, b);
      }
    }
    __syncwarp();                                   // the next rays overwrite W1 / cdf / bins
  }
}

int grid_for_rows(int64_t rows) {
  int64_t blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = (int64_t)mvip_num_sms() * 16;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" {

int mvip_sample_coarse(const float* rays, int ray_stride, int64_t n_rays, const float* t_vals, const float* t_rand,
                       int n_samples, int lindisp, float* z_out, void* stream) {
  MVIP_REQUIRE(ray_stride >= 8 && n_samples >= 1 && n_rays >= 0, MVIP_E_INVALID,
               "mvip_sample_coarse: bad shape (ray_stride=%d n_samples=%d)", ray_stride, n_samples);
  if (n_rays == 0) return MVIP_OK;  // empty batch: nothing to do (pointers may be null)
  MVIP_REQUIRE(rays && t_vals && z_out, MVIP_E_INVALID, "mvip_sample_coarse: null pointer");
  if ((n_samples == 32 || n_samples == 64 || n_samples == 128) && mvip_aligned(z_out, 16) && (!t_rand || mvip_aligned(t_rand, 16))) {
    const int rpw = 128 / n_samples;                 // rays per warp: 4 samples per lane
    int64_t blocks = (n_rays + 8 * rpw - 1) / (8 * rpw);
    const int64_t cap = (int64_t)mvip_num_sms() * 8;       // (16 blocks per SM: the same; uncapped: 43 %)
    if (blocks > cap) blocks = cap;
    auto st = (cudaStream_t)stream;
    if (n_samples == 32) sample_coarse_warp_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(rays, ray_stride, n_rays, t_vals, t_rand, lindisp, z_out);
    else if (n_samples == 64) sample_coarse_warp_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(rays, ray_stride, n_rays, t_vals, t_rand, lindisp, z_out);
    else sample_coarse_warp_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(rays, ray_stride, n_rays, t_vals, t_rand, lindisp, z_out);
    MVIP_LAUNCH_OK("sample_coarse_warp_kernel");
    return MVIP_OK;
  }
  int64_t total = n_rays * n_samples;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  MVIP_REQUIRE(blocks < (1ll << 31), MVIP_E_UNSUPPORTED, "mvip_sample_coarse: too many rays");
  sample_coarse_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(rays, ray_stride, n_rays, t_vals, t_rand,
                                                                               n_samples, lindisp, z_out);
  MVIP_LAUNCH_OK("sample_coarse_kernel");
  return MVIP_OK;
}

int mvip_sample_pdf(const float* bins, const float* weights, const float* u, int u_is_row, int64_t n_rows, int n_bins,
                    int n_out, float* samples, int64_t* inds, float* cdf, void* stream) {
  MVIP_REQUIRE(n_rows >= 0 && n_out >= 1, MVIP_E_INVALID, "mvip_sample_pdf: bad shape");
  MVIP_REQUIRE(n_rows == 0 || (bins && weights && u && samples), MVIP_E_INVALID, "mvip_sample_pdf: null pointer");
  MVIP_REQUIRE(n_bins - 1 >= 8 && n_bins <= kMaxBins && n_out <= kMaxOut, MVIP_E_UNSUPPORTED,
               "mvip_sample_pdf: need 9 <= n_bins <= %d and n_out <= %d (got %d, %d); torch's CPU sum order is only "
               "reproduced for that range", kMaxBins, kMaxOut, n_bins, n_out);
  if (n_rows == 0) return MVIP_OK;
  size_t smem = (size_t)kWarpsPerBlock * 3 * n_bins * sizeof(float);
  sample_pdf_kernel<<<grid_for_rows(n_rows), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, u, u_is_row, n_rows, n_bins, n_out, samples, inds, cdf);
  MVIP_LAUNCH_OK("sample_pdf_kernel");
  return MVIP_OK;
}

int mvip_sample_fine(const float* z_vals, const float* weights, const float* u, int u_is_row, int64_t n_rays,
                     int n_samples, int n_out, float* z_samples, int64_t* inds, float* z_merged, float* z_std,
                     void* stream) {
  MVIP_REQUIRE(n_rays >= 0 && n_out >= 1, MVIP_E_INVALID, "mvip_sample_fine: bad shape");
  MVIP_REQUIRE(n_rays == 0 || (z_vals && weights && u && z_merged), MVIP_E_INVALID, "mvip_sample_fine: null pointer");
  MVIP_REQUIRE(n_samples - 2 >= 8 && n_samples - 1 <= kMaxBins && n_out <= kMaxOut, MVIP_E_UNSUPPORTED,
               "mvip_sample_fine: need 10 <= n_samples <= %d and n_out <= %d (got %d, %d)", kMaxBins + 1, kMaxOut,
               n_samples, n_out);
  if (n_rays == 0) return MVIP_OK;
  if (n_samples == 64 && n_out == 64 && mvip_aligned(z_vals, 16) && mvip_aligned(weights, 16) && mvip_aligned(u, 16) &&
      mvip_aligned(z_merged, 16) && (!z_samples || mvip_aligned(z_samples, 16)) && (!inds || mvip_aligned(inds, 16))) {
    int64_t blocks = (n_rays + 31) / 32;            // 8 warps x 4 rays
    const int64_t cap = (int64_t)mvip_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (u_is_row)
      sample_fine64_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z_vals, weights, u, n_rays, z_samples, inds,
                                                                                     z_merged, z_std);
    else
      sample_fine64_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z_vals, weights, u, n_rays, z_samples, inds,
                                                                                      z_merged, z_std);
    MVIP_LAUNCH_OK("sample_fine64_kernel");
    return MVIP_OK;
  }
  int P2 = 1;
  while (P2 < n_out) P2 <<= 1;
  int B = n_samples - 1;
  size_t per_warp = (size_t)n_samples + 3 * B + P2 + (n_samples + n_out);
  size_t smem = (size_t)kWarpsPerBlock * per_warp * sizeof(float);
  sample_fine_kernel<<<grid_for_rows(n_rays), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      z_vals, weights, u, u_is_row, n_rays, n_samples, n_out, P2, z_samples, inds, z_merged, z_std);
  MVIP_LAUNCH_OK("sample_fine_kernel");
  return MVIP_OK;
}

}  // extern "C"
