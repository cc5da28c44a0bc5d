// embed.cu — standalone positional encoding (the fused MLP kernels compute their own; this serves the
// drop-in Embedder.embed / get_embedder API).   replaces DS_NeRF/run_nerf_helpers.py:22-52
#include "common.cuh"

namespace {
// out[p, :] = [x, sin(x*2^0), cos(x*2^0), ..., sin(x*2^(L-1)), cos(x*2^(L-1))], x = in[p, 0:D]
__global__ void embed_kernel(const float* __restrict__ in, int64_t in_stride, int64_t n, int D, int L,
                             float* __restrict__ out) {
  const int C = D * (1 + 2 * L);
  const int64_t total = n * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / C;
    const int c = (int)(idx - p * C);
    float v;
    if (c < D) {
      v = __ldg(in + p * in_stride + c);
    } else {
      const int t = c - D, k = t / (2 * D), r = t % (2 * D), d = r % D;
      const float x = __fmul_rn(__ldg(in + p * in_stride + d), (float)(1 << k));   // x * freq, freq = 2^k exactly
      v = (r >= D) ? cosf(x) : sinf(x);
    }
    out[idx] = v;
  }
}
}  // namespace

extern "C" int mvip_embed(const float* in, int64_t in_stride, int64_t n, int dims, int num_freqs, float* out,
                          void* stream) {
  MVIP_REQUIRE(n >= 0 && dims >= 1 && dims <= 8 && num_freqs >= 0 && num_freqs <= 24 && in_stride >= dims, MVIP_E_INVALID,
               "mvip_embed: bad shape (dims=%d num_freqs=%d)", dims, num_freqs);
  if (n == 0) return MVIP_OK;
  MVIP_REQUIRE(in && out, MVIP_E_INVALID, "mvip_embed: null pointer");
  const int64_t total = n * dims * (1 + 2 * num_freqs);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)mvip_num_sms() * 32;
  if (blocks > cap) blocks = cap;
  embed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, in_stride, n, dims, num_freqs, out);
  MVIP_LAUNCH_OK("embed_kernel");
  return MVIP_OK;
}
