// Elementwise densities of the fused pass: psi, psi', log-likelihood (density.rs:50-63, 91-103, 122-130).
//
// On B200 DMMA and DFMA issue to the SAME FP64 pipe (profiles/microbench/fp64_pipes_r01.jsonl: mixed kernels are
// additive; ncu on the pass kernel: DMMA sub-pipe % + FP64 pipe % = SM throughput %), so every FP64 instruction
// spent here is taken from the contraction budget (4 N flop per element = 8 pipe-instructions at N = 128).
// Everything below is therefore built to minimise FP64-pipe instructions, moving work to pipes that are idle:
//   * exp(s z) (s = -2 alpha, z = |y| for tanh; s = -alpha/2, z = y^2 for exp): one-fma reduction to |s r| <= ln2/512 (ln2/4096
//     with the BIG tables), the scale s folded into the reduction constants and the series coefficients, a table of 2^(j/N)
//     in shared memory (LSU), the exponent inserted by integer adds (ALU): 8 (7) FP64 instructions, scaling included,
//     instead of 19 for a table-free degree-12 series;
//   * log(1 + e), 1 + e in [1, 2]: 128-entry table {1/v0, -log(1/v0)} indexed by the top mantissa bits (ALU),
//     u = fma(v, 1/v0, -1), |u| <= 2^-8, degree-5 series: 7 FP64 instructions instead of 22;
//   * quotients from an SFU seed (MUFU.RCP64H) and one second-order correction (5 FP64 instructions); |y| and sign
//     transfers by integer ops;
//   * range clamps by integer min on the high word; one e = exp(-2 alpha |y|) shared by tanh, 1 - tanh^2 and log-lik.
// Accuracy: <= ~1e-15 relative per element (the parity bar on the sums G, h, loss is 1e-10; measured ~1e-13).
// Everything is host-callable so the polynomials are checked on the CPU (tests/test_density_host.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "density_tables.inc"

#if defined(__CUDACC__)
#define PICARD_HD __host__ __device__ __forceinline__
#else
#define PICARD_HD inline
#endif

namespace picard {
namespace dmath {

// Two table sets.  SMALL (4 KB) for the kernels that keep W in shared memory; BIG (80 KB) for the row-block kernels, whose
// shared memory only holds the TMA stages: finer tables shorten the polynomials (exp 9 -> 8, log 7 -> 4 FP64 instructions).
template <bool BIG>
struct Tab {
  static constexpr int EXP_BITS = BIG ? 11 : 8;
  static constexpr int LOG_BITS = BIG ? 12 : 7;
  static constexpr int EXP_N = 1 << EXP_BITS;          // doubles
  static constexpr int LOG_N = 2 * ((1 << LOG_BITS) + 1);  // doubles: pairs {1/v0, -log(1/v0)}, one extra pair for v == 2.0
  static constexpr int DOUBLES = EXP_N + LOG_N;
};
constexpr int EXP_TAB_N = Tab<false>::EXP_N;
constexpr int LOG_TAB_N = Tab<false>::LOG_N;
constexpr int TAB_DOUBLES = Tab<false>::DOUBLES;

PICARD_HD int lo32(double t) {
#ifdef __CUDA_ARCH__
  return __double2loint(t);
#else
  uint64_t u; memcpy(&u, &t, 8); return (int)(uint32_t)u;
#endif
}
PICARD_HD int hi32(double t) {
#ifdef __CUDA_ARCH__
  return __double2hiint(t);
#else
  uint64_t u; memcpy(&u, &t, 8); return (int)(uint32_t)(u >> 32);
#endif
}
PICARD_HD double make_double(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double o; memcpy(&o, &u, 8); return o;
#endif
}
PICARD_HD double add_exponent(double p, int k) { return make_double(hi32(p) + (k << 20), lo32(p)); }  // p * 2^k, normal results
// non-negative doubles order like their high words: min(a, limit) on the ALU, not the FP64 pipe (limit given by its high word)
PICARD_HD double clamp_hi(double a_nonneg, int hi_limit) {
  const int h = hi32(a_nonneg);
  return make_double(h < hi_limit ? h : hi_limit, lo32(a_nonneg));
}
// ~20-bit reciprocal seed.  Device: MUFU.RCP64H (SFU pipe, not the FP64 pipe).
PICARD_HD double rcp_seed(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
#else
  // host model of MUFU.RCP64H: only the high word of the operand is read, only the high word of the result is written
  return make_double(hi32(1.0 / make_double(hi32(d), 0)), 0);
#endif
}
// (a / d) for d in a benign range: seed r0 (2^-20), e = 1 - d r0, 1/d = r0 (1 + e + e^2 + O(e^3)), O(e^3) ~ 2^-60.
PICARD_HD double div_seeded(double a, double d) {
  const double r0 = rcp_seed(d);
  const double e = fma(-d, r0, 1.0);
  const double s = fma(e, e, e);
  const double t0 = a * r0;
  return fma(t0, s, t0);
}
// 1/d for d in a benign range (no zero/inf/denormal handling): seed + 2 Newton steps (2^-20 -> 2^-80).
PICARD_HD double rcp_nr(double d) {
  double r = rcp_seed(d);
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Constants of one density, prepared on the host (make_dens_params below).
struct DensParams {
  double alpha, inv_alpha;
  double xscale;   // s: tanh -2 alpha (exponent s |y|) ; exp -alpha / 2 (exponent s y^2)
  int hi_limit;    // high word of the clamp on z = |y| (tanh) or y^2 (exp) that keeps s z >= -700
  // exp(s z) with the scale folded into the range reduction and the series (index 0: SMALL tables, 1: BIG tables)
  double ct[2];    // s N / ln 2
  double nlr[2];   // -ln 2 / (s N)
  double c1, c2, c3, c4;  // s, s^2/2, s^3/6, s^4/24
};

// exp(s z) for z >= 0, s < 0, s z >= -700 (callers clamp z): z = n L + r with L = ln2 / (s N), N = 2^EXP_BITS, |s r| <= ln2/(2N);
// exp(s r) - 1 by a series in r with the powers of s folded into the coefficients, degree 4 (N = 256: (s r)^5/120 <= 4e-17) or
// 3 (N = 2048: (s r)^4/24 <= 4e-17); result T[n mod N] (1 + q) 2^(n div N).  ONE fma for the reduction: the rounding of L
// perturbs the exponent by |s z| 2^-53, a relative error |s z| 1.1e-16 of a result that is <= exp(-|s z|) -- the same order as
// rounding the product s z itself (what a libm-based evaluation of exp(s * z) starts from).  7 FP64 instructions with the BIG
// tables, 8 with the SMALL ones, the multiplication by s included.
template <bool BIG>
PICARD_HD double exp_scaled(double z, const DensParams& dp, const double* __restrict__ T) {
  constexpr int BITS = Tab<BIG>::EXP_BITS, N = 1 << BITS;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() by add/sub, the integer lands in the low word
  const double t = fma(z, dp.ct[BIG ? 1 : 0], MAGIC);
  const double kd = t - MAGIC;
  const int n = lo32(t);
  const double r = fma(kd, dp.nlr[BIG ? 1 : 0], z);
  double p;
  if (BIG) {
    p = fma(r, dp.c3, dp.c2);
  } else {
    p = fma(r, dp.c4, dp.c3);
    p = fma(p, r, dp.c2);
  }
  p = fma(p, r, dp.c1);
  const double q = p * r;
  const double tj = T[n & (N - 1)];
  return add_exponent(fma(tj, q, tj), n >> BITS);
}

// log(v) for v in [1, 2]: i = top LOG_BITS mantissa bits, v0 = 1 + (i + 1/2)/2^LOG_BITS, u = v/v0 - 1 (|u| <= 2^-(LOG_BITS+1)),
// log v = -log(1/v0) + log1p(u) with the series to u^5 (128 entries: truncation u^6/6 <= 6e-16) or u^3 (4096 entries: u^4/4 <= 6e-17).
template <bool BIG>
PICARD_HD double log_1_2(double v, const double* __restrict__ LT) {
  constexpr int BITS = Tab<BIG>::LOG_BITS;
  const int i = (hi32(v) - 0x3FF00000) >> (20 - BITS);  // v == 2.0 exactly (y == 0) lands on the extra entry NI
#ifdef __CUDA_ARCH__
  const double2 rl = reinterpret_cast<const double2*>(LT)[i];
  const double r0 = rl.x, l0 = rl.y;
#else
  const double r0 = LT[2 * i], l0 = LT[2 * i + 1];
#endif
  const double u = fma(v, r0, -1.0);
  if (BIG) {  // u - u^2/2 + u^3/3 = u (1 - u (1/2 - u/3))
    double p = fma(u, -1.0 / 3.0, 0.5);
    p = fma(-u, p, 1.0);
    return fma(p, u, l0);
  }
  double p = fma(u, 0.2, -0.25);
  p = fma(p, u, 1.0 / 3.0);
  p = fma(p, u, -0.5);
  p = p * u;
  return fma(p, u, u) + l0;
}

}  // namespace dmath
using dmath::DensParams;

inline DensParams make_dens_params(int dens, double alpha) {
  DensParams d;
  d.alpha = alpha; d.inv_alpha = 1.0 / alpha;
  const double s = dens == DENS_TANH ? -2.0 * alpha : -0.5 * alpha;
  d.xscale = s;
  const double lim = 700.0 / std::fabs(s);
  d.hi_limit = dmath::hi32(lim);
  const double ln2 = 6.93147180559945309417e-01;
  for (int b = 0; b < 2; ++b) {
    const double n = b ? (double)dmath::Tab<true>::EXP_N : (double)dmath::Tab<false>::EXP_N;
    d.ct[b] = s * n / ln2;
    d.nlr[b] = -ln2 / (s * n);
  }
  d.c1 = s; d.c2 = s * s / 2.0; d.c3 = s * s * s / 6.0; d.c4 = s * s * s * s / 24.0;
  return d;
}

// One element of the density; accumulates the row sums itself so each mode pays only for what it needs.
//   tanh (density.rs:50-63):  psi = tanh(a y), psi' = a (1 - psi^2), loglik = |y| + ln(1 + exp(-2 a |y|)) / a
//   exp  (density.rs:91-103): k = exp(-a y^2 / 2), psi = y k, psi' = (1 - a y^2) k, loglik = -k / a
//   cube (density.rs:122-130): psi = y^3, psi' = 3 y^2, loglik = y^4 / 4
//   linear (internal): psi = y, psi' = 1, loglik = y^2 / 2   (covariance SYRK of the whitening step)
// NEED_PSI: psi / psi' wanted (psi' is added to sd);  NEED_LL: log-likelihood wanted (added to sl).
// tab: [exp table][log table] of the SMALL or BIG set (shared memory on the device).
template <int DENS, bool NEED_PSI, bool NEED_LL, bool BIG = false>
PICARD_HD void density_eval(double y, const DensParams& dp, const double* __restrict__ tab, double& psi, double& psid, double& sd,
                            double& sl) {
  if (DENS == DENS_TANH) {
    const double ay = fabs(y);
    const double e = dmath::exp_scaled<BIG>(dmath::clamp_hi(ay, dp.hi_limit), dp, tab);  // exp(-2 alpha |y|)
    const double v = 1.0 + e;
    if (NEED_PSI) {
      const double th = dmath::div_seeded(1.0 - e, v);  // tanh(alpha |y|)
      psi = copysign(th, y);                            // tanh(alpha y) (alpha > 0): sign transfer on the ALU
      psid = dp.alpha * fma(-th, th, 1.0);
      sd += psid;
    }
    if (NEED_LL) {
      sl += ay;
      sl = fma(dmath::log_1_2<BIG>(v, tab + dmath::Tab<BIG>::EXP_N), dp.inv_alpha, sl);
    }
  } else if (DENS == DENS_EXP) {
    const double y2 = y * y;
    const double k = dmath::exp_scaled<BIG>(dmath::clamp_hi(y2, dp.hi_limit), dp, tab);  // exp(-alpha y^2 / 2)
    if (NEED_PSI) {
      psi = y * k;
      psid = fma(-dp.alpha, y2, 1.0) * k;
      sd += psid;
    }
    if (NEED_LL) sl = fma(-k, dp.inv_alpha, sl);
  } else if (DENS == DENS_CUBE) {
    const double y2 = y * y;
    if (NEED_PSI) {
      psi = y2 * y;
      psid = 3.0 * y2;
      sd += psid;
    }
    if (NEED_LL) sl = fma(0.25 * y2, y2, sl);
  } else {
    if (NEED_PSI) {
      psi = y;
      psid = 1.0;
      sd += 1.0;
    }
    if (NEED_LL) sl = fma(0.5 * y, y, sl);
  }
}

}  // namespace picard
