// Split-precision activations/weights for the tensor-core convolutions ("split" conv mode of lsi.nnutils.nets).
//
// The reference CNN is fp32 throughout (lsi/nnutils/nets.py:263-348); tcgen05 has no fp32 MMA.  A value v is carried as a
// pair of fp16 numbers
//       hi = rn16(v)                    lo = rn16((v - hi) * 2^11)            v ~= hi + lo * 2^-11
// (v - hi is exact in fp32, |v - hi| <= ulp16(hi)/2, so lo is a normal fp16 number of about v's magnitude: no fp16
// subnormals are involved until |v| < 2^-25).  Representation error <= 2^-22 |v|.  A product of two such values is
//       a * w ~= a_hi*w_hi + 2^-11 * (a_hi*w_lo + a_lo*w_hi)                  (dropped a_lo*w_lo term: 2^-22 relative)
// i.e. three kind::f16 MMAs with exact fp16 x fp16 products and fp32 accumulation in TMEM -- 22 mantissa bits per operand
// against TF32's 11 and fp32's 24.  Two accumulators: D0 (hi*hi) and D1 (both cross terms, scaled by 2^11); the epilogue
// forms D0 + 2^-11 * D1.
//
// HBM layout of a split tensor [pixels][C] (C % 32 == 0): per pixel and 32-channel chunk 128 contiguous bytes =
// [hi of the 32 channels | lo of the 32 channels] -- the same byte address (pixel * C + chunk * 32) * 4 and the same total size as
// the fp32 tensor, so one 128B-swizzled TMA box row of 64 fp16 elements is an MMA A-operand row with K = 64: hi | lo.
// Weights are laid out so that ONE MMA sequence over that K produces both accumulators: per (tap, Cout tile) 2N rows,
//       rows [0, N)   = [ w_hi | 0    ]      ->  D0 = a_hi * w_hi
//       rows [N, 2N)  = [ w_lo | w_hi ]      ->  D1 = a_hi * w_lo + a_lo * w_hi
// (N' = 2N columns per MMA; the A tile is read from shared memory four times per chunk, as in the TF32 path, not six).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace lsi {

constexpr float kSplitScale = 2048.f;            // 2^11
constexpr float kSplitInvScale = 1.f / 2048.f;
constexpr float kSplitMax = 65504.f;             // fp16 range: |v| beyond it saturates (activations of this network are O(1..100))

// two values -> packed hi pair, packed lo pair (element 0 in the low half-word)
__device__ __forceinline__ void split_pack2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  v0 = fminf(fmaxf(v0, -kSplitMax), kSplitMax);
  v1 = fminf(fmaxf(v1, -kSplitMax), kSplitMax);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"((v1 - f.y) * kSplitScale), "f"((v0 - f.x) * kSplitScale));
}

__device__ __forceinline__ float2 split_unpack2(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2(fmaf(l.x, kSplitInvScale, h.x), fmaf(l.y, kSplitInvScale, h.y));
}

}  // namespace lsi
