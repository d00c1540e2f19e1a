// picard.hpp -- C++ host-side mirror of the public interface of lmmx/picard-ica v0.1.6 (src/lib.rs:50-60 re-exports) over the
// C ABI of libpicard_b200.so (include/picard_b200.h).  Header-only; link with -lpicard_b200.
//
// Same names, argument meaning and error behaviour as the Rust crate (the reference is compiled code; Rust is not in the
// build image, so this is the compiled-language host mirror -- the Rust `extern "C"` shim itself is shown in INTEGRATION.md):
//   picard_ica::Picard::{fit, fit_with_config, transform}      solver.rs:33,45,199
//   picard_ica::PicardConfig + ConfigBuilder                   config.rs:11-273
//   picard_ica::DensityType::{tanh, tanh_with_alpha, exp, exp_with_alpha, cube}   density.rs:137-176
//   picard_ica::PicardResult::{full_unmixing, mixing}          result.rs:7-64
//   picard_ica::PicardError (InvalidDimensions, SingularMatrix, ComputationError, InvalidConfig)   error.rs:9-42
//   picard_ica::utils::{amari_distance, permute}               utils.rs:16-103
// Matrices are row-major `Array2` (rows, cols, std::vector<double>), like ndarray's standard layout.
// Every computation goes through the CUDA library; there is no CPU fallback (ComputationError without a device).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/picard_b200.h"

namespace picard_ica {

struct Array2 {
  std::size_t rows = 0, cols = 0;
  std::vector<double> data;
  Array2() {}
  Array2(std::size_t r, std::size_t c, double v = 0.0) : rows(r), cols(c), data(r * c, v) {}
  Array2(std::size_t r, std::size_t c, const double* src) : rows(r), cols(c), data(src, src + r * c) {}
  double& operator()(std::size_t i, std::size_t j) { return data[i * cols + j]; }
  double operator()(std::size_t i, std::size_t j) const { return data[i * cols + j]; }
  static Array2 eye(std::size_t n) { Array2 a(n, n); for (std::size_t i = 0; i < n; ++i) a(i, i) = 1.0; return a; }
  Array2 t() const { Array2 o(cols, rows); for (std::size_t i = 0; i < rows; ++i) for (std::size_t j = 0; j < cols; ++j) o(j, i) = (*this)(i, j); return o; }
  Array2 dot(const Array2& b) const {
    if (cols != b.rows) throw std::invalid_argument("Array2::dot: shape mismatch");
    Array2 o(rows, b.cols);
    for (std::size_t i = 0; i < rows; ++i)
      for (std::size_t k = 0; k < cols; ++k) { const double a = (*this)(i, k); for (std::size_t j = 0; j < b.cols; ++j) o(i, j) += a * b(k, j); }
    return o;
  }
};
using Array1 = std::vector<double>;

// ---- error.rs:9-42 ---------------------------------------------------------------------------------------
struct PicardError : std::runtime_error {
  enum class Kind { InvalidDimensions = 1, SingularMatrix = 2, ComputationError = 3, InvalidConfig = 4 };
  Kind kind;
  std::string parameter;  // InvalidConfig only
  PicardError(Kind k, const std::string& display) : std::runtime_error(display), kind(k) {
    if (k == Kind::InvalidConfig) { auto a = display.find('\''); auto b = display.find('\'', a + 1); if (a != std::string::npos && b != std::string::npos) parameter = display.substr(a + 1, b - a - 1); }
  }
};
inline void throw_status(int status, const char* msg) {
  if (status == PICARD_OK) return;
  std::string m = (msg && msg[0]) ? msg : picard_status_string(status);
  throw PicardError(status >= 1 && status <= 4 ? static_cast<PicardError::Kind>(status) : PicardError::Kind::ComputationError, m);
}

// ---- density.rs:137-176 ------------------------------------------------------------------------------------
struct DensityType {
  enum class Kind { Tanh = 0, Exp = 1, Cube = 2 } kind = Kind::Tanh;
  double alpha = 1.0;
  static DensityType tanh() { return {Kind::Tanh, 1.0}; }
  static DensityType tanh_with_alpha(double a) { return {Kind::Tanh, a}; }
  static DensityType exp() { return {Kind::Exp, 1.0}; }
  static DensityType exp_with_alpha(double a) { return {Kind::Exp, a}; }
  static DensityType cube() { return {Kind::Cube, 1.0}; }
};

// ---- config.rs:11-142 ---------------------------------------------------------------------------------------
class ConfigBuilder;
struct PicardConfig {
  DensityType density = DensityType::tanh();
  std::optional<std::size_t> n_components;
  bool ortho = true;
  std::optional<bool> extended;
  bool whiten = true;
  bool centering = true;
  std::size_t max_iter = 500;
  double tol = 1e-7;
  std::size_t m = 7;
  std::size_t ls_tries = 10;
  double lambda_min = 0.01;
  std::optional<Array2> w_init;
  std::optional<std::size_t> fastica_it;
  std::optional<std::size_t> jade_it;
  std::optional<std::uint64_t> random_state;
  bool verbose = false;
  // execution placement (not in the reference)
  int device = -1;
  picard_comm_t* comm = nullptr;
  std::uint32_t flags = 0;

  static PicardConfig new_() { return PicardConfig(); }
  static ConfigBuilder builder();
  bool effective_extended() const { return extended.value_or(ortho); }  // config.rs:99-101
  picard_config_t to_c() const {
    picard_config_t c;
    picard_config_default(&c);
    c.density_kind = static_cast<int>(density.kind); c.alpha = density.alpha;
    c.n_components = n_components ? static_cast<int64_t>(*n_components) : -1;
    c.ortho = ortho; c.extended = extended ? static_cast<int>(*extended) : -1; c.whiten = whiten; c.centering = centering;
    c.max_iter = static_cast<int64_t>(max_iter); c.tol = tol; c.m = static_cast<int64_t>(m); c.ls_tries = static_cast<int64_t>(ls_tries);
    c.lambda_min = lambda_min;
    if (w_init) { c.w_init = w_init->data.data(); c.w_init_rows = static_cast<int64_t>(w_init->rows); c.w_init_cols = static_cast<int64_t>(w_init->cols); }
    c.fastica_it = fastica_it ? static_cast<int64_t>(*fastica_it) : -1;
    c.jade_it = jade_it ? static_cast<int64_t>(*jade_it) : -1;
    c.has_seed = random_state.has_value(); c.seed = random_state.value_or(0);
    c.verbose = verbose; c.device = device; c.comm = comm; c.flags = flags;
    return c;
  }
  void validate() const {  // config.rs:104-142
    picard_config_t c = to_c();
    char err[512];
    throw_status(picard_config_validate(&c, err, sizeof err), err);
  }
};
class ConfigBuilder {  // config.rs:145-273
  PicardConfig c_;
 public:
  ConfigBuilder& density(DensityType d) { c_.density = d; return *this; }
  ConfigBuilder& n_components(std::size_t n) { c_.n_components = n; return *this; }
  ConfigBuilder& ortho(bool v) { c_.ortho = v; return *this; }
  ConfigBuilder& extended(bool v) { c_.extended = v; return *this; }
  ConfigBuilder& whiten(bool v) { c_.whiten = v; return *this; }
  ConfigBuilder& centering(bool v) { c_.centering = v; return *this; }
  ConfigBuilder& max_iter(std::size_t v) { c_.max_iter = v; return *this; }
  ConfigBuilder& tol(double v) { c_.tol = v; return *this; }
  ConfigBuilder& m(std::size_t v) { c_.m = v; return *this; }
  ConfigBuilder& ls_tries(std::size_t v) { c_.ls_tries = v; return *this; }
  ConfigBuilder& lambda_min(double v) { c_.lambda_min = v; return *this; }
  ConfigBuilder& w_init(const Array2& w) { c_.w_init = w; return *this; }
  ConfigBuilder& fastica_it(std::size_t v) { c_.fastica_it = v; return *this; }
  ConfigBuilder& jade_it(std::size_t v) { c_.jade_it = v; return *this; }
  ConfigBuilder& random_state(std::uint64_t v) { c_.random_state = v; return *this; }
  ConfigBuilder& verbose(bool v) { c_.verbose = v; return *this; }
  ConfigBuilder& device(int v) { c_.device = v; return *this; }
  PicardConfig build() const { return c_; }
  PicardConfig build_validated() const { c_.validate(); return c_; }
};
inline ConfigBuilder PicardConfig::builder() { return ConfigBuilder(); }

// ---- result.rs:7-129 ------------------------------------------------------------------------------------------
struct PicardResult {
  std::optional<Array2> whitening;
  Array2 unmixing;
  Array2 sources;
  std::optional<Array1> mean;
  std::size_t n_iterations = 0;
  bool converged = false;
  double gradient_norm = 0.0;
  std::optional<Array1> signs;
  picard_stats_t stats{};

  Array2 full_unmixing() const { return whitening ? unmixing.dot(*whitening) : unmixing; }  // result.rs:39-44
  Array2 mixing() const {  // result.rs:49-64: (W^T W)^-1 W^T by Gauss-Jordan, transpose fallback
    const Array2 w = full_unmixing(), wt = w.t(), wtw = wt.dot(w);
    const std::size_t n = wtw.rows;
    Array2 aug(n, 2 * n);
    for (std::size_t i = 0; i < n; ++i) { for (std::size_t j = 0; j < n; ++j) aug(i, j) = wtw(i, j); aug(i, n + i) = 1.0; }
    for (std::size_t i = 0; i < n; ++i) {
      std::size_t mr = i;
      for (std::size_t k = i + 1; k < n; ++k) if (std::fabs(aug(k, i)) > std::fabs(aug(mr, i))) mr = k;
      for (std::size_t j = 0; j < 2 * n; ++j) std::swap(aug(i, j), aug(mr, j));
      if (std::fabs(aug(i, i)) < 1e-15) return wt;
      const double piv = aug(i, i);
      for (std::size_t j = 0; j < 2 * n; ++j) aug(i, j) /= piv;
      for (std::size_t k = 0; k < n; ++k) if (k != i) { const double f = aug(k, i); for (std::size_t j = 0; j < 2 * n; ++j) aug(k, j) -= f * aug(i, j); }
    }
    Array2 inv(n, n);
    for (std::size_t i = 0; i < n; ++i) for (std::size_t j = 0; j < n; ++j) inv(i, j) = aug(i, n + j);
    return inv.dot(wt);
  }
};

// ---- solver.rs:23-215 -------------------------------------------------------------------------------------------
struct Picard {
  static PicardResult fit(const Array2& x) { return fit_with_config(x, PicardConfig()); }  // solver.rs:33
  static PicardResult fit_with_config(const Array2& x, const PicardConfig& config) {       // solver.rs:45
    config.validate();  // solver.rs:46: before looking at the data
    picard_config_t c = config.to_c();
    picard_result_t r;
    char err[1024];
    const int st = picard_fit(x.data.empty() ? nullptr : x.data.data(), static_cast<int64_t>(x.rows), static_cast<int64_t>(x.cols),
                              static_cast<int64_t>(x.cols), &c, &r, err, sizeof err);
    throw_status(st, err);
    PicardResult out;
    const std::size_t nc = static_cast<std::size_t>(r.n_components), nf = static_cast<std::size_t>(r.n_features), t = static_cast<std::size_t>(r.n_samples);
    if (r.whitening) out.whitening = Array2(nc, nf, r.whitening);
    out.unmixing = Array2(nc, nc, r.unmixing);
    if (r.sources) out.sources = Array2(nc, t, r.sources);
    if (r.mean) out.mean = Array1(r.mean, r.mean + nf);
    out.n_iterations = static_cast<std::size_t>(r.n_iterations); out.converged = r.converged != 0; out.gradient_norm = r.gradient_norm;
    if (r.signs) out.signs = Array1(r.signs, r.signs + nc);
    out.stats = r.stats;
    picard_result_free(&r);
    return out;
  }
  static Array2 transform(const Array2& x, const PicardResult& result, int device = -1) {  // solver.rs:199
    picard_result_t r{};
    r.n_components = static_cast<int64_t>(result.unmixing.rows);
    r.n_features = static_cast<int64_t>(result.whitening ? result.whitening->cols : result.unmixing.rows);
    r.whitening = result.whitening ? const_cast<double*>(result.whitening->data.data()) : nullptr;
    r.unmixing = const_cast<double*>(result.unmixing.data.data());
    r.mean = result.mean ? const_cast<double*>(result.mean->data()) : nullptr;
    Array2 out(result.unmixing.rows, x.cols);
    char err[1024];
    throw_status(picard_transform(x.data.empty() ? nullptr : x.data.data(), static_cast<int64_t>(x.rows), static_cast<int64_t>(x.cols),
                                  static_cast<int64_t>(x.cols), &r, out.data.data(), device, err, sizeof err), err);
    return out;
  }
};

// ---- utils.rs:16-103 (N x N host code in the reference too) -----------------------------------------------------
namespace utils {
inline double amari_distance(const Array2& w, const Array2& a) {  // utils.rs:82-103
  Array2 p = w.dot(a);
  for (double& v : p.data) v = std::fabs(v);
  auto s = [](const Array2& r) {
    double sum = 0.0;
    for (std::size_t i = 0; i < r.rows; ++i) {
      double rs = 0.0, rm = 0.0;
      for (std::size_t j = 0; j < r.cols; ++j) { const double q = r(i, j) * r(i, j); rs += q; rm = std::max(rm, q); }
      if (rm > 1e-15) sum += rs / rm - 1.0;
    }
    return sum;
  };
  return (s(p) + s(p.t())) / (2.0 * static_cast<double>(p.rows));
}
inline Array2 permute(const Array2& a_in, bool scale) {  // utils.rs:16-68
  Array2 a = a_in;
  const std::size_t n = a.rows;
  bool done = false;
  while (!done) {
    done = true;
    for (std::size_t i = 0; i < n; ++i)
      for (std::size_t j = 0; j < i; ++j)
        if (a(i, i) * a(i, i) + a(j, j) * a(j, j) < a(i, j) * a(i, j) + a(j, i) * a(j, i)) {
          for (std::size_t c = 0; c < a.cols; ++c) std::swap(a(i, c), a(j, c));
          done = false;
        }
  }
  if (scale)
    for (std::size_t i = 0; i < n; ++i) { const double d = a(i, i); if (std::fabs(d) > 1e-10) for (std::size_t j = 0; j < a.cols; ++j) a(i, j) /= d; }
  std::vector<double> cs(n, 0.0);
  for (std::size_t j = 0; j < n; ++j) for (std::size_t i = 0; i < n; ++i) cs[j] += std::fabs(a(i, j));
  std::vector<std::size_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](std::size_t x, std::size_t y) { return cs[x] < cs[y]; });
  Array2 r(n, n);
  for (std::size_t i = 0; i < n; ++i) for (std::size_t j = 0; j < n; ++j) r(i, j) = a(order[i], order[j]);
  return r;
}
}  // namespace utils

}  // namespace picard_ica
