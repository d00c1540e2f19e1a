// Elementwise densities of the fused pass: psi, psi', log-likelihood (density.rs:50-63, 91-103, 122-130).
//
// On B200 DMMA and DFMA issue to the SAME FP64 pipe (profiles/microbench/fp64_pipes_r01.jsonl: mixed
// kernels are additive), so every FP64 instruction spent here is taken from the contraction budget.
// The transcendental kernels below therefore (a) share one e = exp(-2 alpha |y|) between tanh, 1 - tanh^2
// and the log-likelihood, (b) use argument ranges known a priori (x <= 0; 1 + e in (1, 2]) to drop all
// special-case handling, (c) take reciprocal seeds from the SFU (MUFU.RCP64H), which is a separate pipe.
// Accuracy target: <= ~1e-14 relative per element (the parity bar on G, h, loss is 1e-10).
// Everything is __host__ __device__ so the polynomials can be checked on the CPU (tests/test_density_host).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "common.cuh"

#if defined(__CUDACC__)
#define PICARD_HD __host__ __device__ __forceinline__
#else
#define PICARD_HD inline
#endif

namespace picard {
namespace dmath {

PICARD_HD int lo32(double t) {
#ifdef __CUDA_ARCH__
  return __double2loint(t);
#else
  uint64_t u; memcpy(&u, &t, 8); return (int)(uint32_t)u;
#endif
}
PICARD_HD double add_exponent(double p, int k) {  // p * 2^k for normal results
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  uint64_t u; memcpy(&u, &p, 8); u += (uint64_t)((int64_t)k << 52); double o; memcpy(&o, &u, 8); return o;
#endif
}
// ~20-bit reciprocal seed.  Device: MUFU.RCP64H (SFU pipe, not the FP64 pipe).
PICARD_HD double rcp_seed(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
#else
  return (double)(1.0f / (float)d);
#endif
}
// 1/d for d in a benign range (no zero/inf/denormal handling): seed + 2 Newton steps (quadratic: 2^-20 -> 2^-80).
PICARD_HD double rcp_nr(double d) {
  double r = rcp_seed(d);
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// exp(x) for x <= ~0 (any x in [-700, 700] works); x below -700 returns exp(-700) ~ 1e-304 (never matters:
// it is only ever added to 1).  Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-12 Taylor
// (truncation <= 1.7e-16 relative at the interval ends), exponent inserted by integer add.
PICARD_HD double exp_nonpos(double x) {
  x = fmax(x, -700.0);
  const double L2E = 1.4426950408889634074, LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() by add/sub, integer lands in the low word
  double t = fma(x, L2E, MAGIC);
  double kd = t - MAGIC;
  int k = lo32(t);
  double r = fma(kd, -LN2HI, x);
  r = fma(kd, -LN2LO, r);
  double p = 2.08767569878680989792e-09;            // 1/12!
  p = fma(p, r, 2.50521083854417187751e-08);        // 1/11!
  p = fma(p, r, 2.75573192239858906526e-07);        // 1/10!
  p = fma(p, r, 2.75573192239858906526e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873015873e-05);        // 1/8!
  p = fma(p, r, 1.98412698412698412698e-04);        // 1/7!
  p = fma(p, r, 1.38888888888888888889e-03);        // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);        // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);        // 1/4!
  p = fma(p, r, 1.66666666666666666667e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return add_exponent(p, k);
}

// log(1 + e) for e in [0, 1]:  1+e in [1,2]; fold to [sqrt(1/2), sqrt 2] by an optional halving, then
// log1p(f) = 2 atanh(s), s = f / (2 + f), |s| <= 0.1716, series to s^17 (truncation <= 9e-16 relative).
PICARD_HD double log1p_unit(double e) {
  const double SQRT2M1 = 0.41421356237309504880, LN2 = 0.69314718055994530942;
  bool big = e > SQRT2M1;
  double f = big ? fma(e, 0.5, -0.5) : e;
  double d = 2.0 + f;
  double s = f * rcp_nr(d);
  double z = s * s;
  double p = 2.0 / 17.0;
  p = fma(p, z, 2.0 / 15.0);
  p = fma(p, z, 2.0 / 13.0);
  p = fma(p, z, 2.0 / 11.0);
  p = fma(p, z, 2.0 / 9.0);
  p = fma(p, z, 2.0 / 7.0);
  p = fma(p, z, 2.0 / 5.0);
  p = fma(p, z, 2.0 / 3.0);
  double res = fma(s * z, p, s + s);
  return big ? res + LN2 : res;
}

}  // namespace dmath

// One element of the density.  NEED_PSI: psi and psi' wanted; NEED_LL: log-likelihood wanted.
//   tanh (density.rs:50-63):  psi = tanh(a y), psi' = a (1 - psi^2), loglik = |y| + ln(1 + exp(-2 a |y|)) / a
//   exp  (density.rs:91-103): k = exp(-a y^2 / 2), psi = y k, psi' = (1 - a y^2) k, loglik = -k / a
//   cube (density.rs:122-130): psi = y^3, psi' = 3 y^2, loglik = y^4 / 4
//   linear (internal): psi = y, psi' = 1, loglik = y^2 / 2   (covariance SYRK of the whitening step)
template <int DENS, bool NEED_PSI, bool NEED_LL>
PICARD_HD void density_eval(double y, double alpha, double inv_alpha, double& psi, double& psid, double& ll) {
  if (DENS == DENS_TANH) {
    double ay = fabs(y);
    double e = dmath::exp_nonpos(-2.0 * alpha * ay);
    if (NEED_PSI) {
      double r = dmath::rcp_nr(1.0 + e);
      double th = (1.0 - e) * r;                  // tanh(alpha |y|)
      psi = copysign(th, y * alpha);              // tanh(alpha y)
      psid = alpha * fma(-th, th, 1.0);
    }
    if (NEED_LL) ll = fma(dmath::log1p_unit(e), inv_alpha, ay);
  } else if (DENS == DENS_EXP) {
    double y2 = y * y;
    double k = dmath::exp_nonpos(-0.5 * alpha * y2);
    if (NEED_PSI) {
      psi = y * k;
      psid = fma(-alpha, y2, 1.0) * k;
    }
    if (NEED_LL) ll = -k * inv_alpha;
  } else if (DENS == DENS_CUBE) {
    double y2 = y * y;
    if (NEED_PSI) {
      psi = y2 * y;
      psid = 3.0 * y2;
    }
    if (NEED_LL) ll = 0.25 * (y2 * y2);
  } else {
    if (NEED_PSI) {
      psi = y;
      psid = 1.0;
    }
    if (NEED_LL) ll = 0.5 * y * y;
  }
}

}  // namespace picard
