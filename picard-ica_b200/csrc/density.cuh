// Elementwise densities of the fused pass: psi, psi', log-likelihood (density.rs:50-63, 91-103, 122-130).
//
// On B200 DMMA and DFMA issue to the SAME FP64 pipe (profiles/microbench/fp64_pipes_r01.jsonl: mixed kernels are
// additive; ncu on the pass kernel: DMMA sub-pipe % + FP64 pipe % = SM throughput %), so every FP64 instruction
// spent here is taken from the contraction budget (4 N flop per element = 8 pipe-instructions at N = 128).
// Everything below is therefore built to minimise FP64-pipe instructions, moving work to pipes that are idle:
//   * exp: Cody-Waite reduction to |r| <= ln2/512 with a 256-entry table of 2^(j/256) in shared memory (LSU) and
//     the exponent inserted by integer adds (ALU): 9 FP64 instructions instead of 18 for a table-free degree-12 series;
//   * log(1 + e), 1 + e in [1, 2]: 128-entry table {1/v0, -log(1/v0)} indexed by the top mantissa bits (ALU),
//     u = fma(v, 1/v0, -1), |u| <= 2^-8, degree-5 series: 7 FP64 instructions instead of 22;
//   * reciprocal seeds from the SFU (MUFU.RCP64H) + 2 Newton steps; |y| and sign transfers by integer ops;
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
  static constexpr int LOG_N = 2 << LOG_BITS;          // doubles: pairs {1/v0, -log(1/v0)}
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
  return (double)(1.0f / (float)d);
#endif
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

// exp(x) for x in [-700, 700] (callers clamp): x = n ln2/N + r, N = 2^EXP_BITS, |r| <= ln2/(2N); exp(r) - 1 by a series of
// degree 4 (N = 256: r^5/120 <= 4e-17) or 3 (N = 2048: r^4/24 <= 4e-17); e = T[n mod N] (1 + q) 2^(n div N).
template <bool BIG>
PICARD_HD double exp_tab(double x, const double* __restrict__ T) {
  constexpr int BITS = Tab<BIG>::EXP_BITS, N = 1 << BITS;
  const double C = 1.4426950408889634074 * N;  // N / ln 2
  const double LHI = 6.93147180369123816490e-01 / N, LLO = 1.90821492927058770002e-10 / N;  // ln2/N split; LHI has 32 trailing zero bits
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() by add/sub, the integer lands in the low word
  const double t = fma(x, C, MAGIC);
  const double kd = t - MAGIC;
  const int n = lo32(t);
  double r = fma(kd, -LHI, x);
  r = fma(kd, -LLO, r);
  double p;
  if (BIG) {
    p = fma(r, 1.0 / 6.0, 0.5);
  } else {
    p = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    p = fma(p, r, 0.5);
  }
  p = fma(p, r, 1.0);
  const double q = p * r;
  const double tj = T[n & (N - 1)];
  return add_exponent(fma(tj, q, tj), n >> BITS);
}

// log(v) for v in [1, 2]: i = top LOG_BITS mantissa bits, v0 = 1 + (i + 1/2)/2^LOG_BITS, u = v/v0 - 1 (|u| <= 2^-(LOG_BITS+1)),
// log v = -log(1/v0) + log1p(u) with the series to u^5 (128 entries: truncation u^6/6 <= 6e-16) or u^3 (4096 entries: u^4/4 <= 6e-17).
template <bool BIG>
PICARD_HD double log_1_2(double v, const double* __restrict__ LT) {
  constexpr int BITS = Tab<BIG>::LOG_BITS, NI = 1 << BITS;
  int i = (hi32(v) - 0x3FF00000) >> (20 - BITS);
  i = i < NI - 1 ? i : NI - 1;  // v == 2.0 exactly
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

// Constants of one density, prepared on the host.
struct DensParams {
  double alpha, inv_alpha;
  double xscale;   // tanh: -2 alpha (x = xscale * |y|) ; exp: -alpha / 2 (x = xscale * y^2)
  int hi_limit;    // high word of the clamp on |y| (tanh) or y^2 (exp) that keeps x >= -700
};
inline DensParams make_dens_params(int dens, double alpha) {
  DensParams d;
  d.alpha = alpha; d.inv_alpha = 1.0 / alpha;
  d.xscale = dens == DENS_TANH ? -2.0 * alpha : -0.5 * alpha;
  const double lim = 700.0 / std::fabs(d.xscale);
  d.hi_limit = dmath::hi32(lim);
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
    const double x = dmath::clamp_hi(ay, dp.hi_limit) * dp.xscale;
    const double e = dmath::exp_tab<BIG>(x, tab);
    const double v = 1.0 + e;
    if (NEED_PSI) {
      const double r = dmath::rcp_nr(v);
      const double th = (1.0 - e) * r;            // tanh(alpha |y|)
      psi = copysign(th, y * dp.alpha);           // tanh(alpha y): sign transfer on the ALU (alpha > 0 in practice)
      psid = dp.alpha * fma(-th, th, 1.0);
      sd += psid;
    }
    if (NEED_LL) {
      sl += ay;
      sl = fma(dmath::log_1_2<BIG>(v, tab + dmath::Tab<BIG>::EXP_N), dp.inv_alpha, sl);
    }
  } else if (DENS == DENS_EXP) {
    const double y2 = y * y;
    const double x = dmath::clamp_hi(y2, dp.hi_limit) * dp.xscale;
    const double k = dmath::exp_tab<BIG>(x, tab);
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
