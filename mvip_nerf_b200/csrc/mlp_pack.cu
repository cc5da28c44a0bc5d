// mlp_pack.cu — fp32 nn.Parameters -> bf16 pre-swizzled weight blob streamed by the MLP kernels.
// The blob is laid out in consumption order so the TMA producer issues plain 1-D bulk copies.
#include "mlp_common.cuh"

namespace {
using namespace mlp;

struct Src {
  int param;      // index into ParamPtrs
  int rows;       // n extent of the chunk image (256 / 128)
  int base;       // element offset of (n=0,k=0)
  int stride_n;   // element stride along n
  int stride_k;   // element stride along k
  int valid_k;    // columns >= valid_k are zero padding
};

__device__ __forceinline__ Src fwd_src(int c) {
  if (c == 0) return {kPW(0), 256, 0, 63, 1, 63};
  if (c <= 16) return {kPW(1 + (c - 1) / 4), 256, 64 * ((c - 1) % 4), 256, 1, 64};
  if (c == 17) return {kPW(5), 256, 0, 319, 1, 63};
  if (c <= 21) return {kPW(5), 256, 63 + 64 * (c - 18), 319, 1, 64};
  if (c <= 29) return {kPW(6 + (c - 22) / 4), 256, 64 * ((c - 22) % 4), 256, 1, 64};
  if (c <= 33) return {kPFeatW, 256, 64 * (c - 30), 256, 1, 64};
  if (c <= 37) return {kPViewsW, 128, 64 * (c - 34), 283, 1, 64};
  return {kPViewsW, 128, 256, 283, 1, 27};
}
// dgrad operand: element (n = input feature, k = output feature) = W[row_off + k][col_off + n]
__device__ __forceinline__ Src bwd_src(int c) {
  if (c <= 1) return {kPViewsW, 256, 64 * c * 283, 1, 283, 64};
  if (c <= 5) return {kPFeatW, 256, 64 * (c - 2) * 256, 1, 256, 64};
  int layer = 7 - (c - 6) / 4, j = (c - 6) % 4;
  int ld = (layer == 5) ? 319 : 256;
  return {kPW(layer), 256, 64 * j * ld + ((layer == 5) ? 63 : 0), 1, ld, 64};
}

__global__ void pack_kernel(ParamPtrs pp, uint8_t* __restrict__ packed) {
  const int total_groups = (int)((kFwdBytes + kBwdBytes) / 16);
  for (int gidx = blockIdx.x * blockDim.x + threadIdx.x; gidx < total_groups; gidx += gridDim.x * blockDim.x) {
    size_t byte = (size_t)gidx * 16;
    Src s;
    size_t chunk_base;
    uint32_t in_chunk;
    if (byte < kFwdBytes) {
      int c = byte < 34 * (size_t)kW256 ? (int)(byte / kW256) : 34 + (int)((byte - 34 * (size_t)kW256) / kW128);
      s = fwd_src(c);
      chunk_base = fwd_chunk_off(c);
      in_chunk = (uint32_t)(byte - chunk_base);
    } else {
      size_t b2 = byte - kFwdBytes;
      int c = (int)(b2 / kW256);
      s = bwd_src(c);
      chunk_base = kFwdBytes + (size_t)c * kW256;
      in_chunk = (uint32_t)(b2 - (size_t)c * kW256);
    }
    // invert chunk_off16: in_chunk = (n>>3)*1024 + (n&7)*128 + ((g ^ (n&7))<<4)
    uint32_t n = (in_chunk >> 10) * 8 + ((in_chunk >> 7) & 7);
    uint32_t g = ((in_chunk >> 4) & 7) ^ (n & 7);
    const float* src = pp.p[s.param] + s.base + (size_t)n * s.stride_n;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int k = g * 8 + e;
      v[e] = (k < s.valid_k) ? __ldg(src + (size_t)k * s.stride_k) : 0.f;
    }
    uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(packed + chunk_base + in_chunk) = o;
  }
  // fp32 tail
  float* sm = reinterpret_cast<float*>(packed + kSmallOff);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kSmallFloats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < kSmBiasFeat) v = pp.p[2 * (i / 256) + 1][i % 256];
    else if (i < kSmBiasViews) v = pp.p[kPFeatB][i - kSmBiasFeat];
    else if (i < kSmWAlpha) v = pp.p[kPViewsB][i - kSmBiasViews];
    else if (i < kSmBAlpha) v = pp.p[kPAlphaW][i - kSmWAlpha];
    else if (i == kSmBAlpha) v = pp.p[kPAlphaB][0];
    else if (i >= kSmWRgb && i < kSmBRgb) v = pp.p[kPRgbW][i - kSmWRgb];
    else if (i >= kSmBRgb && i < kSmBRgb + 3) v = pp.p[kPRgbB][i - kSmBRgb];
    sm[i] = v;
  }
}

}  // namespace

extern "C" {

size_t mvip_mlp_packed_bytes(void) { return mlp::kPackedBytes; }

int mvip_mlp_pack_weights(const float* const* params, void* packed, void* stream) {
  MVIP_REQUIRE(params && packed, MVIP_E_INVALID, "mvip_mlp_pack_weights: null pointer");
  MVIP_REQUIRE(mvip_aligned(packed, 1024), MVIP_E_INVALID, "mvip_mlp_pack_weights: packed must be 1024-byte aligned");
  mlp::ParamPtrs pp;
  for (int i = 0; i < MVIP_MLP_NUM_PARAMS; ++i) {
    MVIP_REQUIRE(params[i], MVIP_E_INVALID, "mvip_mlp_pack_weights: params[%d] is null", i);
    pp.p[i] = params[i];
  }
  pack_kernel<<<mvip_num_sms() * 2, 256, 0, (cudaStream_t)stream>>>(pp, static_cast<uint8_t*>(packed));
  MVIP_LAUNCH_OK("pack_kernel");
  return MVIP_OK;
}

}  // extern "C"
