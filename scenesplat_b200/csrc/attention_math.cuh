// Scalar helpers shared by the attention forward and backward kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ss {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = i + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.5e-5, 26x below the bf16 rounding of P), i added into the exponent field.
// Needs x <= 126; anything below -126 (masked keys: -inf) comes out as 2^-126 * 0.99993 (a denormal: nothing next
// to the row maximum's weight of >= 2^-8; the clamp keeps the exponent-field addition from borrowing into the sign).
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517166769f, 0.24261112209f);
  p = fmaf(p, f, 0.69326098571f);
  p = fmaf(p, f, 0.99992807355f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
// The same without the clamp, for arguments known to lie in [-126, 126].
__device__ __forceinline__ float exp2_poly_bounded(float x) {
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517166769f, 0.24261112209f);
  p = fmaf(p, f, 0.69326098571f);
  p = fmaf(p, f, 0.99992807355f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}


}  // namespace ss
