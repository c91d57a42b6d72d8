// FP16 pair operand format of the tensor-core kernels (tc_gemm.cu, tc_conv.cu) and of the kernels that produce their
// operands (bn.cu, vocr_split_f16_f32):   x * 2^e  ->  hi = fp16(x 2^e),  lo = fp16((x 2^e - hi) * 2^11)
// e is a per-tensor power of two that puts an upper bound of max|x| into [2^14, 2^15): 22 significant bits wherever hi
// is a normal FP16 number, and an absolute error floor of bound * 2^-50 below that (lo keeps >= 11 bits down to 2^-25
// after scaling) - i.e. a 22-bit format with ~40 binades of dynamic range under the bound, so even a bound that is
// 2^10 too loose costs nothing measurable.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace vocr {

constexpr float kPairLoScale = 2048.f;  // lo plane holds (x - hi) * 2^11

// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }
// x * 2^sh for any sh in [-252, 252] without intermediate overflow of the multiplier
__device__ __forceinline__ float scale_pow2(float x, int sh) {
  const int s1 = max(-126, min(126, sh));
  return x * exp2i(s1) * exp2i(sh - s1);
}
// exponent for a tensor whose magnitudes are bounded by the (non-negative) float with these bits
__device__ __forceinline__ int pair_exponent(unsigned bound_bits) {
  const int e_field = (int)((bound_bits >> 23) & 0xffu);  // bound < 2^(e_field - 126)
  return max(-126, min(126, 141 - e_field));              // bound * 2^e < 2^15
}
// four consecutive elements (already multiplied by 2^e) -> 8 bytes of the hi plane and 8 bytes of the lo plane
__device__ __forceinline__ void pair_pack4(float a0, float a1, float a2, float a3, uint2& ph, uint2& pl) {
  const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((a0 - f01.x) * kPairLoScale, (a1 - f01.y) * kPairLoScale);
  const __half2 l23 = __floats2half2_rn((a2 - f23.x) * kPairLoScale, (a3 - f23.y) * kPairLoScale);
  ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
  pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
}

}  // namespace vocr
