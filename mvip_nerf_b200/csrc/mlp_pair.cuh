// mlp_pair.cuh — device helpers shared by the CTA-pair (cta_group::2) kernels whose activations live in tensor memory
// (forward: mlp_forward.cu, dgrad chain: mlp_backward.cu).
#pragma once
#include "mlp_common.cuh"

namespace mlp {

constexpr uint32_t kSlotBytes2 = 64 * 128;     // weight ring slot: 64 weight rows x 64 k (bf16) = this CTA's part of one K chunk of an N-half
constexpr int kGroupBars2 = 4;                 // ring of (full, empty) barrier pairs, one per (layer, N-half) group of slots

// UMMA smem descriptor (SWIZZLE_128B, K-major, LBO 16 B, SBO 1024 B) split into its two words, so that the issuer
// only adds to the low word: lo = (addr >> 4) | (1 << 16), hi = 64 | version 1 (bit 14) | layout 2 (bits 29..31)
constexpr uint32_t kDescHi2 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo2(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void mma2_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi2)
      : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi2)
      : "memory");
}
// cta_group::1, both operands in shared memory, descriptor low words only (wgrad)
__device__ __forceinline__ void mma1_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi2)
      : "memory");
}
// MN-major SWIZZLE_128B operand: LBO = stride between 64-wide MN blocks, SBO = 1024 (8 K rows)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// This thread's 64 accumulator columns -> registers (four loads in flight, one round trip).
__device__ __forceinline__ void load_half(uint32_t tD, uint32_t (&raw)[4][16]) {
  tmem_ld16(tD, raw[0]);
  tmem_ld16(tD + 16, raw[1]);
  tmem_ld16(tD + 32, raw[2]);
  tmem_ld16(tD + 48, raw[3]);
  tmem_ld_wait_on16(raw[0]);
  tmem_ld_wait_on16(raw[1]);
  tmem_ld_wait_on16(raw[2]);
  tmem_ld_wait_on16(raw[3]);
}

// packed bf16 pair -> both halves kept / zeroed by two mask bits (bit 0: low half, bit 16: high half of `sel`)
__device__ __forceinline__ uint32_t mask_bf16x2(uint32_t pk, uint32_t sel) { return pk & (sel * 0xFFFFu); }

}  // namespace mlp
