// Compiles against picard-ica_b200/host/picard.hpp and libpicard_b200.so.  Mode "cpu": config defaults / validation / error
// mapping / utils, and -- on a box without a GPU -- the loud ComputationError of fit.  Mode "gpu": the reference's own solver
// tests (solver.rs:288-408) through the C++ mirror.  Prints "ok <n checks>" or exits non-zero.
#include <cstdio>
#include <cstring>
#include <random>

#include "../../picard-ica_b200/host/picard.hpp"

using namespace picard_ica;
static int checks = 0;
#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); return 1; } ++checks; } while (0)

static Array2 laplace_mixture(std::size_t n, std::size_t t, unsigned seed) {  // like generate_test_data, solver.rs:257-286
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> u(1e-6, 1.0);
  std::normal_distribution<double> nd;
  Array2 s(n, t), a(n, n);
  for (auto& v : s.data) { const double m = -std::log(u(g)); v = (u(g) < 0.5) ? -m : m; }
  for (auto& v : a.data) v = nd(g);
  return a.dot(s);
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && !strcmp(argv[1], "gpu");
  // config.rs defaults and validation
  PicardConfig d;
  CHECK(d.ortho && !d.extended && d.whiten && d.centering && d.max_iter == 500 && d.tol == 1e-7 && d.m == 7 && d.ls_tries == 10 && d.lambda_min == 0.01);
  CHECK(d.effective_extended());
  try { PicardConfig::builder().max_iter(0).build_validated(); CHECK(false); }
  catch (const PicardError& e) { CHECK(e.kind == PicardError::Kind::InvalidConfig && e.parameter == "max_iter"); }
  try { PicardConfig::builder().fastica_it(5).jade_it(5).build_validated(); CHECK(false); }  // solver.rs:397-407
  catch (const PicardError& e) { CHECK(e.kind == PicardError::Kind::InvalidConfig && e.parameter == "jade_it"); }
  // utils.rs:147-171
  Array2 a(2, 2); a(0, 0) = 1; a(0, 1) = 0.5; a(1, 0) = 0.3; a(1, 1) = 1;
  const double det = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
  Array2 inv(2, 2); inv(0, 0) = a(1, 1) / det; inv(0, 1) = -a(0, 1) / det; inv(1, 0) = -a(1, 0) / det; inv(1, 1) = a(0, 0) / det;
  CHECK(utils::amari_distance(inv, a) < 1e-10);
  Array2 pm(2, 2); pm(0, 0) = 0.1; pm(0, 1) = 0.9; pm(1, 0) = 0.95; pm(1, 1) = 0.05;
  Array2 pp = utils::permute(pm, true);
  CHECK(std::fabs(pp(0, 0) - 1.0) < 1e-6 && std::fabs(pp(1, 1) - 1.0) < 1e-6);
  if (!gpu) {
    if (picard_device_count() == 0) {
      try { Picard::fit(laplace_mixture(3, 100, 1)); CHECK(false); }
      catch (const PicardError& e) { CHECK(e.kind == PicardError::Kind::ComputationError); }
    }
    try { Picard::fit(Array2()); CHECK(false); }
    catch (const PicardError& e) { CHECK(e.kind == PicardError::Kind::InvalidDimensions || e.kind == PicardError::Kind::ComputationError); }
    printf("ok %d\n", checks);
    return 0;
  }
  // solver.rs:289-301 test_fit_default
  Array2 x = laplace_mixture(3, 1000, 42);
  PicardResult r = Picard::fit(x);
  CHECK(r.unmixing.rows == 3 && r.unmixing.cols == 3 && r.sources.rows == 3 && r.sources.cols == 1000 && r.whitening && r.mean);
  // solver.rs:304-318 test_fit_with_config
  PicardResult r2 = Picard::fit_with_config(laplace_mixture(4, 2000, 42), PicardConfig::builder().n_components(3).max_iter(100).tol(1e-6).random_state(42).build());
  CHECK(r2.unmixing.rows == 3 && r2.sources.cols == 2000 && r2.whitening->rows == 3 && r2.whitening->cols == 4 && r2.n_iterations <= 100);
  // solver.rs:359-375 test_transform
  Array2 y = Picard::transform(x, r);
  CHECK(y.rows == 3 && y.cols == 1000);
  double md = 0; for (std::size_t i = 0; i < y.data.size(); ++i) md = std::max(md, std::fabs(y.data[i] - r.sources.data[i]));
  CHECK(md < 1e-9);
  // solver.rs:378-394 test_no_whiten
  PicardResult r3 = Picard::fit_with_config(laplace_mixture(3, 1000, 9), PicardConfig::builder().whiten(false).max_iter(50).random_state(3).build());
  CHECK(!r3.whitening && r3.unmixing.rows == 3);
  // jade / fastica warm starts (solver.rs:321-356)
  PicardResult r4 = Picard::fit_with_config(laplace_mixture(4, 3000, 5), PicardConfig::builder().jade_it(10).max_iter(100).build());
  CHECK(r4.unmixing.rows == 4 && (r4.converged || r4.gradient_norm < 1.0));
  PicardResult r5 = Picard::fit_with_config(laplace_mixture(4, 3000, 5), PicardConfig::builder().fastica_it(5).max_iter(100).random_state(1).build());
  CHECK(r5.converged);
  // separation quality: W_full A close to a scaled permutation; mixing() is its pseudo-inverse
  Array2 m = r.mixing(), wm = r.full_unmixing().dot(m);
  for (std::size_t i = 0; i < 3; ++i) for (std::size_t j = 0; j < 3; ++j) CHECK(std::fabs(wm(i, j) - (i == j ? 1.0 : 0.0)) < 1e-8);
  try { Picard::fit_with_config(x, PicardConfig::builder().w_init(Array2::eye(2)).build()); CHECK(false); }  // solver.rs:100-108
  catch (const PicardError& e) { CHECK(e.kind == PicardError::Kind::InvalidDimensions); }
  printf("ok %d\n", checks);
  return 0;
}
