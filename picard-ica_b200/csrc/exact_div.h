// term / k of the Taylor series of matrix_exp (math.rs:60) without the generic double-precision division sequence.
// Host-callable so that tests/host/div_by_count_check.c can run the very same arithmetic on the CPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PICARD_EXACT_DIV_HD __host__ __device__ __forceinline__
#else
#define PICARD_EXACT_DIV_HD static inline
#endif

// a / kk for a small positive integer kk (2 .. 30 here) and rcp = RN(1 / kk): q = RN(a rcp), r = a - kk q (exact: FMA),
// q + r rcp rounded once is the correctly rounded quotient (Markstein's correction step).  Zeros keep their sign through a * rcp;
// operands near the ends of the exponent range (denormal quotients, infinities, NaN) take the plain division.
PICARD_EXACT_DIV_HD double div_by_count(double a, double kk, double rcp) {
  const double q = a * rcp;
  const double aa = fabs(a);
  if (aa > 1e-270 && aa < 1e270) return fma(fma(-kk, q, a), rcp, q);
  if (aa == 0.0) return q;
  return a / kk;
}
