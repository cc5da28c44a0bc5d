// umma_selftest.cu — one-tile tcgen05 GEMMs that pin the descriptor / shared-memory layout
// conventions the MLP kernels rely on (K-major operands for forward + dgrad, MN-major for wgrad).
// Exposed as mvip_selftest_umma (include/mvip_nerf.h); exercised by tests/test_gpu_umma.py.
#include "common.cuh"

namespace {

// which = 0: A [128,K] and B [N,K] row-major fp32 (K contiguous)  -> K-major chunk images
// which = 1: A [K,128] and B [K,N] row-major fp32 (M/N contiguous) -> MN-major chunk images,
//            LBO = stride between 64-wide MN blocks, SBO = 1024 (8 K-rows)
// which = 2: as 1 with the LBO/SBO roles swapped (diagnostic only)
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int which, const float* __restrict__ a, const float* __restrict__ b, int N, int K,
                     float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  uint8_t* sA;
  uint8_t* sB;
  uint32_t a_img_bytes, b_img_bytes;
  if (which == 0) {
    // chunk image c of A: rows = 128 (M), 64 k-columns [64c, 64c+64)
    a_img_bytes = 128 * 128;
    b_img_bytes = N * 128;
    sA = smem;
    sB = smem + (K / 64) * a_img_bytes;
    for (int idx = tid; idx < 128 * (K / 8); idx += 128) {
      int r = idx / (K / 8), g8 = idx % (K / 8);
      int c = g8 / 8, g = g8 % 8;
      const float* src = a + (size_t)r * K + g8 * 8;
      uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]),
                           pack_bf16x2(src[6], src[7]));
      *reinterpret_cast<uint4*>(sA + c * a_img_bytes + chunk_off16(r, g)) = v;
    }
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
      int r = idx / (K / 8), g8 = idx % (K / 8);
      int c = g8 / 8, g = g8 % 8;
      const float* src = b + (size_t)r * K + g8 * 8;
      uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]),
                           pack_bf16x2(src[6], src[7]));
      *reinterpret_cast<uint4*>(sB + c * b_img_bytes + chunk_off16(r, g)) = v;
    }
  } else {
    // chunk image c of A: rows = K (points), 64 m-columns [64c, 64c+64)
    a_img_bytes = K * 128;
    b_img_bytes = K * 128;
    sA = smem;
    sB = smem + 2 * a_img_bytes;
    for (int idx = tid; idx < K * (128 / 8); idx += 128) {
      int r = idx / 16, g8 = idx % 16;
      int c = g8 / 8, g = g8 % 8;
      const float* src = a + (size_t)r * 128 + g8 * 8;
      uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]),
                           pack_bf16x2(src[6], src[7]));
      *reinterpret_cast<uint4*>(sA + c * a_img_bytes + chunk_off16(r, g)) = v;
    }
    for (int idx = tid; idx < K * (N / 8); idx += 128) {
      int r = idx / (N / 8), g8 = idx % (N / 8);
      int c = g8 / 8, g = g8 % 8;
      const float* src = b + (size_t)r * N + g8 * 8;
      uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]),
                           pack_bf16x2(src[6], src[7]));
      *reinterpret_cast<uint4*>(sB + c * b_img_bytes + chunk_off16(r, g)) = v;
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    if (which == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      for (int ks = 0; ks < K / 16; ++ks) {
        int c = ks / 4, kk = ks % 4;
        uint64_t da = umma_desc_sw128(smem_u32(sA + c * a_img_bytes) + kk * 32, 16, 1024);
        uint64_t db = umma_desc_sw128(smem_u32(sB + c * b_img_bytes) + kk * 32, 16, 1024);
        umma_bf16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
      }
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);
      for (int ks = 0; ks < K / 16; ++ks) {
        uint32_t lbo_a = a_img_bytes, lbo_b = b_img_bytes, sbo = 1024;
        uint64_t da, db;
        if (which == 1) {
          da = umma_desc_sw128(smem_u32(sA) + ks * 2048, lbo_a, sbo);
          db = umma_desc_sw128(smem_u32(sB) + ks * 2048, lbo_b, sbo);
        } else {
          da = umma_desc_sw128(smem_u32(sA) + ks * 2048, sbo, lbo_a);
          db = umma_desc_sw128(smem_u32(sB) + ks * 2048, sbo, lbo_b);
        }
        umma_bf16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace

extern "C" int mvip_selftest_umma(int which, const float* a, const float* b, int N, int K, float* out, void* stream) {
  MVIP_REQUIRE(a && b && out, MVIP_E_INVALID, "mvip_selftest_umma: null pointer");
  MVIP_REQUIRE(which >= 0 && which <= 2, MVIP_E_INVALID, "mvip_selftest_umma: which=%d", which);
  MVIP_REQUIRE((N == 64 || N == 128 || N == 256) && K >= 64 && K <= 256 && K % 64 == 0, MVIP_E_UNSUPPORTED,
               "mvip_selftest_umma: N=%d K=%d unsupported", N, K);
  size_t bytes = (which == 0) ? (size_t)(K / 64) * (128 * 128 + N * 128) : (size_t)(2 + N / 64) * K * 128;
  bytes += 1024;
  MVIP_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  umma_selftest_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(which, a, b, N, K, out);
  MVIP_LAUNCH_OK("umma_selftest_kernel");
  return MVIP_OK;
}
