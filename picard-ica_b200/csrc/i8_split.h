// i8_split.h -- the error-free digit splitting behind the INT8 tensor-core passes, as plain host/device C++ (no CUDA headers):
// the kernels (i8_common.cuh) and the host check (tests/host/i8_split_check.cpp) compile the same code.  See i8_common.cuh.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PICARD_I8_HD __host__ __device__ inline
#else
#define PICARD_I8_HD inline
#endif

namespace picard {
namespace i8 {

constexpr int S = 6;                          // balanced radix-256 digits per operand
constexpr int NPROD = S * (S + 1) / 2;        // slice products with p + q <= S - 1
constexpr int FRAC_BITS = 8 * S - 1;          // I = rint(v 2^(FRAC_BITS - e))
constexpr uint64_t DIGIT_BIAS = 0x808080808080ull;  // sum_{p < 6} 128 256^p
static_assert(S == 6, "DIGIT_BIAS and the level combination below are written for six digits");

// exponent e with 1.008 m < 2^e for the maximum magnitude m of a row / sample (0 for an all-zero one): the largest balanced
// six-digit integer is 127 (256^6 - 1) / 255 = 0.99608 2^47, so |v| <= m gives a representable I = rint(v 2^(47 - e))
PICARD_I8_HD int bound_exponent(double m) { return m > 0.0 ? ilogb(m * 1.008) + 1 : 0; }

// the S digits of v (|v| <= m, e = bound_exponent(m)) as the low S bytes of the result, most significant digit in byte S-1, two's complement
PICARD_I8_HD uint64_t split_digits(double v, int e) {
#ifdef __CUDA_ARCH__
  const long long I = __double2ll_rn(scalbn(v, FRAC_BITS - e));
#else
  const long long I = (long long)std::nearbyint(std::scalbn(v, FRAC_BITS - e));
#endif
  return ((uint64_t)I + DIGIT_BIAS) ^ DIGIT_BIAS;  // DIGIT_BIAS is also the mask of the digits' top bits
}
PICARD_I8_HD int digit_of(uint64_t digits, int p) { return (int)(int8_t)(uint8_t)(digits >> (8 * (S - 1 - p))); }

// the value of the six level sums (s32, exact) in units of 2^(ea + eb - 30): hi + 2^-24 lo, rounded once
PICARD_I8_HD double combine_levels(int l0, int l1, int l2, int l3, int l4, int l5) {
  const long long hi = (long long)l0 * 65536 + (long long)l1 * 256 + (long long)l2;
  const long long lo = (long long)l3 * 65536 + (long long)l4 * 256 + (long long)l5;
  return fma((double)lo, 5.9604644775390625e-08 /* 2^-24 */, (double)hi);
}
constexpr int COMBINE_EXP = -30;  // 2^-14 of the two digit scalings and 2^-16 of the level combination

}  // namespace i8
}  // namespace picard
