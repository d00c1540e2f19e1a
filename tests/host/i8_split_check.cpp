// Host check of the error-free digit splitting behind the INT8 tensor-core passes (picard-ica_b200/csrc/i8_split.h -- PRODUCT
// code, compiled here with g++).  Emulates exactly what the kernels do with integers: digits by split_digits(), the slice products
// with p + q <= S - 1 accumulated per level in int32 (overflow checked), combine_levels(), power-of-two scaling; and compares
// with long double / __float128-free exact references built from the same inputs.
//   (1) digit round trip;  (2) the LOSS contraction y = sum_k w_k x_k (K = 128; per-row / per-sample exponents);
//   (3) the gradient contraction g = sum_t psi_t y_t over T samples with FIXED exponents (psi: tanh bound, y: Cauchy-Schwarz row bound)
//       including the flush schedule (every 16384 samples);  (4) adversarial ranges.
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../picard-ica_b200/csrc/i8_split.h"

using namespace picard::i8;

static double value_of_digits(uint64_t dg, int e) {  // exact in long double for 48-bit integers
  long double s = 0;
  for (int p = 0; p < S; ++p) s += (long double)digit_of(dg, p) * ldexpl(1.0L, e - 8 * p - 7);
  return (double)s;
}

struct Levels {
  long long l[S] = {0, 0, 0, 0, 0, 0};
  bool overflow = false;
  void add(uint64_t a, uint64_t b) {
    for (int p = 0; p < S; ++p)
      for (int q = 0; p + q < S; ++q) l[p + q] += (long long)digit_of(a, p) * digit_of(b, q);
  }
  void check() { for (int d = 0; d < S; ++d) if (std::llabs(l[d]) >= (1ll << 31)) overflow = true; }
  double value(int ea, int eb) {
    check();
    return std::ldexp(combine_levels((int)l[0], (int)l[1], (int)l[2], (int)l[3], (int)l[4], (int)l[5]), ea + eb + COMBINE_EXP);
  }
  void clear() { for (int d = 0; d < S; ++d) l[d] = 0; }
};

int main() {
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nrm;
  std::uniform_real_distribution<double> uni(-1.0, 1.0);
  auto laplace = [&]() { double u = uni(rng) * 0.5; return -std::copysign(1.0, u) * std::log(1 - 2 * std::fabs(u)) / std::sqrt(2.0); };

  // (1) round trip: |v - digits| <= 2^(e - 48) (half a quantum), digits in [-128, 127], top digit never saturates past the bound
  double rt = 0;
  for (int i = 0; i < 2000000; ++i) {
    const double m = std::ldexp(1.0 + std::fabs(uni(rng)), (int)(uni(rng) * 40));
    const int e = bound_exponent(m);
    const double v = (i % 3 == 0) ? std::copysign(m, uni(rng)) : m * uni(rng);  // includes the bound itself
    const double back = value_of_digits(split_digits(v, e), e);
    rt = std::fmax(rt, std::fabs(back - v) / std::ldexp(1.0, e - FRAC_BITS));
  }

  // (2) LOSS contraction, K = 128
  const int K = 128;
  double loss_err = 0, loss_err_adv = 0;
  bool ovf = false;
  for (int trial = 0; trial < 4000; ++trial) {
    const bool adv = trial >= 2000;  // adversarial: W' row with entries spread over 1e8, one component of the sample 1e6 x larger
    std::vector<double> w(K), x(K);
    double mw = 0, mx = 0;
    for (int k = 0; k < K; ++k) {
      w[k] = nrm(rng) / std::sqrt((double)K) * (adv ? std::pow(10.0, -8.0 * std::fabs(uni(rng))) : 1.0);
      x[k] = (k & 1) ? laplace() : uni(rng) * 1.7;
      if (adv && k == 17) x[k] *= 1e6;
      mw = std::fmax(mw, std::fabs(w[k])); mx = std::fmax(mx, std::fabs(x[k]));
    }
    const int ew = bound_exponent(mw), ex = bound_exponent(mx);
    Levels L;
    long double ref = 0, scale = 0;
    for (int k = 0; k < K; ++k) {
      L.add(split_digits(w[k], ew), split_digits(x[k], ex));
      ref += (long double)w[k] * x[k];
      scale += std::fabs((long double)w[k] * x[k]);
    }
    const double got = L.value(ew, ex);
    ovf |= L.overflow;
    // error relative to max|w| max|x| (what the splitting controls); the kernels' outputs are compared as max|dy| / max|y|
    const double err = std::fabs((double)(got - ref)) / (mw * mx);
    if (adv) loss_err_adv = std::fmax(loss_err_adv, err); else loss_err = std::fmax(loss_err, err);
  }

  // (3) gradient contraction over T samples, fixed exponents, flush every 16384 samples
  const int T = 200000;
  double grad_err = 0, grad_diag_err = 0;
  for (int pair = 0; pair < 6; ++pair) {
    const bool diag = pair < 3;
    // |w_j| |x_t|_max bound: whitened N = 128 data has |x_t| up to ~20, rows of an orthogonal W have norm 1
    const int ey = bound_exponent(20.0), ep = bound_exponent(1.0);
    Levels L;
    long double ref = 0;
    double g = 0;
    for (int t = 0; t < T; ++t) {
      const double yi = (pair & 1) ? laplace() : uni(rng) * 1.7320508;
      const double yj = diag ? yi : laplace();
      const double psi = std::tanh(yi);
      L.add(split_digits(yj, ey), split_digits(psi, ep));
      ref += (long double)psi * yj;
      if ((t + 1) % 16384 == 0 || t + 1 == T) { g += L.value(ey, ep); ovf |= L.overflow; L.clear(); }
    }
    // relative to max|G| ~ T E[psi(y) y] (the diagonal): the tolerance definition of the parity tests
    const double err = std::fabs((double)(g - ref)) / (0.5 * T);
    if (diag) grad_diag_err = std::fmax(grad_diag_err, err); else grad_err = std::fmax(grad_err, err);
  }
  printf("{\"round_trip_quanta\": %.3f, \"loss_err\": %.3e, \"loss_err_adversarial\": %.3e, \"grad_err\": %.3e, \"grad_diag_err\": %.3e, "
         "\"overflow\": %d}\n", rt, loss_err, loss_err_adv, grad_err, grad_diag_err, ovf ? 1 : 0);
  return 0;
}
