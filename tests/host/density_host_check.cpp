// Host-side accuracy check of the table-driven exp / log / tanh / log-lik kernels in picard-ica_b200/csrc/density.cuh
// (compiled as plain C++: the same source the device runs).  Reference: long double libm (64-bit mantissa).
// Prints one JSON line: max relative errors.  Used by tests/test_density_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../picard-ica_b200/csrc/density.cuh"

using namespace picard;

static const double EXP_TAB_S[] = PICARD_EXP_TAB_INIT;
static const double LOG_TAB_S[] = PICARD_LOG_TAB_INIT;
static const double EXP_TAB_B[] = PICARD_EXP_TAB_BIG_INIT;
static const double LOG_TAB_B[] = PICARD_LOG_TAB_BIG_INIT;

template <bool BIG>
static int run() {
  using TB = dmath::Tab<BIG>;
  static double tab[TB::DOUBLES];
  for (int i = 0; i < TB::EXP_N; ++i) tab[i] = BIG ? EXP_TAB_B[i] : EXP_TAB_S[i];
  for (int i = 0; i < TB::LOG_N; ++i) tab[TB::EXP_N + i] = BIG ? LOG_TAB_B[i] : LOG_TAB_S[i];
  double max_exp = 0, max_log = 0, max_psi = 0, max_psid_abs = 0, max_ll = 0, max_k = 0;
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  const int n = 2000000;
  for (int it = 0; it < n; ++it) {
    // exp(s z) on s z in [-700, 0] for several scales s; the error is measured in units of (3 + |s z|) 2^-53: the one-fma
    // reduction perturbs the exponent by |s z| 2^-53 (density.cuh), the same order as rounding the product s z
    {
      static const double alphas[4] = {1.0, 0.1, 0.7, 2.5};
      const DensParams dpe = make_dens_params(it & 1, alphas[(it >> 1) & 3]);
      const double sc = dpe.xscale;
      double x = (it % 3 == 0) ? -700.0 * rnd() : ((it % 3 == 1) ? -40.0 * rnd() : -2.0 * rnd());
      double z = x / sc;
      if (it == 0) z = 0.0;
      if (sc * z < -700.0) z = -700.0 / sc * (1 - 1e-15);
      double e = dmath::exp_scaled<BIG>(z, dpe, tab);
      long double xe = (long double)sc * (long double)z;
      long double er = expl(xe);
      double rel = (double)(fabsl(((long double)e - er) / er) / ((3.0L + fabsl(xe)) * 1.1102230246251565e-16L));
      if (rel > max_exp) max_exp = rel;
    }
    // log on [1, 2]
    double v = 1.0 + rnd();
    if (it == 0) v = 1.0;
    if (it == 1) v = 2.0;
    double l = dmath::log_1_2<BIG>(v, tab + TB::EXP_N);
    long double lr = logl((long double)v);
    double ab = (double)fabsl((long double)l - lr);
    if (ab > max_log) max_log = ab;
  }
  for (int dens = 0; dens < 2; ++dens) {
    for (double alpha : {1.0, 0.1, 0.7, 2.5}) {
      DensParams dp = make_dens_params(dens, alpha);
      for (int it = 0; it < 400000; ++it) {
        double y = (it % 4 == 0) ? 60.0 * (rnd() - 0.5) : ((it % 4 == 1) ? 8.0 * (rnd() - 0.5) : ((it % 4 == 2) ? 1e-3 * (rnd() - 0.5) : 2000.0 * (rnd() - 0.5)));
        if (it == 0) y = 0.0;
        if (it == 1) y = -0.0;
        if (it == 2) y = 1e300;
        double psi = 0, psid = 0, sd = 0, sl = 0;
        if (dens == 0) {
          density_eval<DENS_TANH, true, true, BIG>(y, dp, tab, psi, psid, sd, sl);
          long double a = alpha, yy = y;
          long double ps = tanhl(a * yy), pd = a * (1 - ps * ps), ll = fabsl(yy) + logl(1 + expl(-2 * a * fabsl(yy))) / a;
          double r1 = (double)fabsl(((long double)psi - ps));  // absolute: (1 - e) cancels for tiny |y|, like any exp-based tanh
          if (r1 > max_psi) max_psi = r1;
          double r2 = (double)fabsl((long double)psid - pd);
          if (r2 > max_psid_abs) max_psid_abs = r2;
          double r3 = (double)(fabsl((long double)sl - ll) / ll);
          if (r3 > max_ll) max_ll = r3;
        } else {
          if (fabs(y) > 100) continue;
          density_eval<DENS_EXP, true, true, BIG>(y, dp, tab, psi, psid, sd, sl);
          long double a = alpha, yy = y;
          long double k = expl(-a * yy * yy / 2);
          double r1 = (double)fabsl((long double)psi - yy * k), r2 = (double)fabsl((long double)psid - (1 - a * yy * yy) * k);
          double r3 = (double)fabsl((long double)sl + k / a) * alpha;
          double m = fmax(r1 / (fabs(y) + 1.0), fmax(r2 / (1 + alpha * y * y), r3));  // absolute, scaled by the prefactor only
          if (m > max_k) max_k = m;
        }
      }
    }
  }
  printf("{\"set\": \"%s\", \"exp_err_units\": %.3e, \"log_abs\": %.3e, \"tanh_psi_abs\": %.3e, \"tanh_psid_abs\": %.3e, \"tanh_ll_rel\": %.3e, \"expdens_abs\": %.3e}\n",
         BIG ? "big" : "small", max_exp, max_log, max_psi, max_psid_abs, max_ll, max_k);
  return 0;
}

int main() { return run<false>() | run<true>(); }
