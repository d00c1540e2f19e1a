// =====================================================================================================
// picard_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the fit path of lmmx/picard-ica v0.1.6 (Rust, /root/reference/src/*.rs), written
// from the reference's behaviour, function by function, each citing the file:line it follows.  It exists
// to CHECK the CUDA product (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
// legs).  The product (picard-ica_b200/) never imports, links or executes anything in this directory.
//
// PARITY PIN STATUS: the reference cannot be compiled here (no cargo/rustc) and its own tests hold no
// golden vectors for G, h, loss, the iterate sequence, whitening K, JADE or the final unmixing
// (SURVEY.md §4, §8c).  For those quantities this oracle is "PARITY UNPINNED": it is validated by
//   (1) the reference's own known-answer tests re-expressed (math.rs:100-152, utils.rs:146-208,
//       whitening.rs:123-150, jade.rs:208-256, lbfgs.rs:178-202) -- tests/test_oracle_known_answers.py;
//   (2) an independent numpy restatement (oracle/numpy_ref.py) agreeing to rounding;
//   (3) invariants: finite-difference of the loss vs the projected relative gradient, Amari -> 0.
// The N x N helpers (sln_det, sym_decorrelation, matrix_exp(0), skew, amari) ARE pinned by (1).
//
// Third-party arithmetic the reference reaches that is not under /root/reference:
//   ndarray 0.17.1 `.dot` -> cblas_dgemm ; ndarray-linalg 0.18.0 / lax 0.18.0 -> dgesvd, dsyev, dgetrf ;
//   openblas-src 0.10.13 (system OpenBLAS).  This oracle calls the SAME routines from the OpenBLAS that
//   ships in this image (scipy.libs/libscipy_openblas-*.so, LP64, `scipy_`-prefixed symbols).
//   rand 0.9.2 StdRng (ChaCha12) + rand_distr 0.5.1 StandardNormal are NOT reproduced: the random
//   w_init path uses the build's own documented generator (splitmix64 + Box-Muller, see orc_rng below);
//   parity runs always pass w_init explicitly (config.rs:226-229; used verbatim, solver.rs:98-111).
//
// Cost structure is kept like the reference's on purpose (serial elementwise passes, fresh N x T
// temporaries, threaded BLAS only inside dgemm/dgesvd), because this file is also the CPU baseline.
// =====================================================================================================
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>
#include <limits>
#include <string>
#include <utility>
#include <vector>

extern "C" {
void scipy_cblas_dgemm(int order, int ta, int tb, int m, int n, int k, double alpha, const double* a, int lda,
                       const double* b, int ldb, double beta, double* c, int ldc);
void scipy_dgesvd_(const char* jobu, const char* jobvt, const int* m, const int* n, double* a, const int* lda,
                   double* s, double* u, const int* ldu, double* vt, const int* ldvt, double* work,
                   const int* lwork, int* info, size_t, size_t);
void scipy_dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w,
                  double* work, const int* lwork, int* info, size_t, size_t);
void scipy_dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {

enum { CblasRowMajor = 101, CblasNoTrans = 111, CblasTrans = 112 };

// status codes shared with include/picard_b200.h (error.rs:9-42)
enum { ORC_OK = 0, ORC_INVALID_DIMENSIONS = 1, ORC_SINGULAR = 2, ORC_COMPUTATION = 3, ORC_INVALID_CONFIG = 4 };
enum { DENS_TANH = 0, DENS_EXP = 1, DENS_CUBE = 2 };

// ---- a minimal row-major matrix that does NOT zero-fill on allocation (ndarray's mapv/dot allocate and
// write once; a zero-filling container would add a pass the reference does not make)
struct Mat {
  int64_t r = 0, c = 0;
  double* p = nullptr;
  Mat() {}
  Mat(int64_t r_, int64_t c_) : r(r_), c(c_) { p = (double*)malloc(sizeof(double) * (size_t)(r * c > 0 ? r * c : 1)); }
  Mat(const Mat& o) : r(o.r), c(o.c) {
    p = (double*)malloc(sizeof(double) * (size_t)(r * c > 0 ? r * c : 1));
    memcpy(p, o.p, sizeof(double) * (size_t)(r * c));
  }
  Mat(Mat&& o) noexcept : r(o.r), c(o.c), p(o.p) { o.p = nullptr; o.r = o.c = 0; }
  Mat& operator=(Mat o) { std::swap(r, o.r); std::swap(c, o.c); std::swap(p, o.p); return *this; }
  ~Mat() { free(p); }
  double& operator()(int64_t i, int64_t j) { return p[i * c + j]; }
  double operator()(int64_t i, int64_t j) const { return p[i * c + j]; }
  int64_t size() const { return r * c; }
  static Mat zeros(int64_t r, int64_t c) { Mat m(r, c); for (int64_t i = 0; i < r * c; ++i) m.p[i] = 0.0; return m; }
  static Mat eye(int64_t n) { Mat m = zeros(n, n); for (int64_t i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
  static Mat from(const double* src, int64_t r, int64_t c) { Mat m(r, c); memcpy(m.p, src, sizeof(double) * (size_t)(r * c)); return m; }
};

// C = op(A) * op(B), row-major, via cblas_dgemm -- what ndarray's `.dot` lowers to with the blas feature.
Mat dot(const Mat& a, bool ta, const Mat& b, bool tb) {
  int64_t m = ta ? a.c : a.r, k = ta ? a.r : a.c, n = tb ? b.r : b.c;
  Mat c(m, n);
  if (m == 0 || n == 0) return c;
  if (k == 0) { for (int64_t i = 0; i < m * n; ++i) c.p[i] = 0; return c; }
  scipy_cblas_dgemm(CblasRowMajor, ta ? CblasTrans : CblasNoTrans, tb ? CblasTrans : CblasNoTrans, (int)m, (int)n, (int)k,
                    1.0, a.p, (int)a.c, b.p, (int)b.c, 0.0, c.p, (int)n);
  return c;
}

inline double rust_signum(double v) {  // f64::signum: +0.0 -> 1, -0.0 -> -1, NaN -> NaN   (quirk Q8)
  if (std::isnan(v)) return v;
  return std::signbit(v) ? -1.0 : 1.0;
}
inline double max_abs(const Mat& a) {  // iter().map(abs).fold(0.0, f64::max): NaN entries are ignored
  double m = 0.0;
  for (int64_t i = 0; i < a.size(); ++i) m = std::fmax(m, std::fabs(a.p[i]));
  return m;
}

// Rust's `{:.4e}`: "1.2345e-3" (no zero padding, no '+')
std::string rust_e4(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  char buf[64];
  snprintf(buf, sizeof buf, "%.4e", v);
  std::string s(buf);
  size_t e = s.find('e');
  std::string mant = s.substr(0, e);
  int ex = atoi(s.c_str() + e + 1);
  return mant + "e" + std::to_string(ex);
}

// ------------------------------------------------------------------------------------------------
// density.rs
// ------------------------------------------------------------------------------------------------
// log_lik: density.rs:50-56 (tanh; uses ln(1+x), not ln_1p -- quirk Q12), 91-94 (exp), 122-124 (cube)
void log_lik(int kind, double alpha, const double* y, int64_t n, double* out) {
  if (kind == DENS_TANH) {
    for (int64_t i = 0; i < n; ++i) {
      double a = std::fabs(y[i]);
      out[i] = a + std::log(1.0 + std::exp(-2.0 * alpha * a)) / alpha;
    }
  } else if (kind == DENS_EXP) {
    for (int64_t i = 0; i < n; ++i) out[i] = -std::exp(-alpha * y[i] * y[i] / 2.0) / alpha;
  } else {
    for (int64_t i = 0; i < n; ++i) { double v2 = y[i] * y[i]; out[i] = (v2 * v2) / 4.0; }
  }
}
// score_and_der: density.rs:58-63 (tanh), 96-103 (exp: four passes with temporaries), 126-130 (cube)
void score_and_der(int kind, double alpha, const Mat& y, Mat& psi, Mat& psid) {
  int64_t n = y.size();
  psi = Mat(y.r, y.c);
  psid = Mat(y.r, y.c);
  if (kind == DENS_TANH) {
    for (int64_t i = 0; i < n; ++i) psi.p[i] = std::tanh(alpha * y.p[i]);
    for (int64_t i = 0; i < n; ++i) psid.p[i] = alpha * (1.0 - psi.p[i] * psi.p[i]);
  } else if (kind == DENS_EXP) {
    Mat ysq(y.r, y.c), k(y.r, y.c);
    for (int64_t i = 0; i < n; ++i) ysq.p[i] = y.p[i] * y.p[i];
    for (int64_t i = 0; i < n; ++i) k.p[i] = std::exp(-alpha / 2.0 * ysq.p[i]);
    for (int64_t i = 0; i < n; ++i) psi.p[i] = y.p[i] * k.p[i];
    for (int64_t i = 0; i < n; ++i) psid.p[i] = (1.0 - alpha * ysq.p[i]) * k.p[i];
  } else {
    for (int64_t i = 0; i < n; ++i) psi.p[i] = y.p[i] * y.p[i] * y.p[i];
    for (int64_t i = 0; i < n; ++i) psid.p[i] = 3.0 * y.p[i] * y.p[i];
  }
}

// ------------------------------------------------------------------------------------------------
// math.rs
// ------------------------------------------------------------------------------------------------
// sln_det: math.rs:84-88 -> ndarray-linalg Determinant::sln_det -> LAPACK dgetrf.  An exactly singular
// factorisation (info > 0) is reported by ndarray-linalg as (0, -inf), not as an error.
int sln_det(const Mat& m, double* sign, double* logabs) {
  int n = (int)m.r, info = 0;
  if (n == 0) { *sign = 1.0; *logabs = 0.0; return ORC_OK; }
  Mat a(m);  // LU in place on a copy (row-major memory read as the transpose: same determinant)
  std::vector<int> piv(n);
  scipy_dgetrf_(&n, &n, a.p, &n, piv.data(), &info);
  if (info < 0) return ORC_COMPUTATION;
  if (info > 0) { *sign = 0.0; *logabs = -std::numeric_limits<double>::infinity(); return ORC_OK; }
  double s = 1.0, l = 0.0;
  for (int i = 0; i < n; ++i) {
    if (piv[i] != i + 1) s = -s;
    double u = a(i, i);
    if (u < 0) s = -s;
    if (u == 0) { *sign = 0.0; *logabs = -std::numeric_limits<double>::infinity(); return ORC_OK; }
    l += std::log(std::fabs(u));
  }
  *sign = s; *logabs = l;
  return ORC_OK;
}

// eigh(UPLO::Lower) of a symmetric row-major matrix -> ascending eigenvalues, eigenvectors in columns.
int eigh_lower(const Mat& a, std::vector<double>& w, Mat& v) {
  int n = (int)a.r, info = 0, lwork = -1;
  // row-major lower triangle == column-major upper triangle of the same symmetric matrix
  Mat cm(a);
  w.assign(n, 0.0);
  double wq = 0;
  scipy_dsyev_("V", "U", &n, cm.p, &n, w.data(), &wq, &lwork, &info, 1, 1);
  lwork = (int)wq;
  std::vector<double> work((size_t)(lwork > 1 ? lwork : 1));
  scipy_dsyev_("V", "U", &n, cm.p, &n, w.data(), work.data(), &lwork, &info, 1, 1);
  if (info != 0) return ORC_COMPUTATION;
  v = Mat(n, n);  // cm is column-major: eigenvector j is cm[j*n .. j*n+n)
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) v(i, j) = cm.p[(size_t)j * n + i];
  return ORC_OK;
}

// sym_decorrelation: math.rs:12-33.   W <- (W W^T)^{-1/2} W ; min eigenvalue < 1e-10 -> SingularMatrix
int sym_decorrelation(const Mat& w, Mat& out) {
  Mat wwt = dot(w, false, w, true);
  std::vector<double> ev; Mat u;
  if (eigh_lower(wwt, ev, u) != ORC_OK) return ORC_COMPUTATION;
  double mn = std::numeric_limits<double>::infinity();
  for (double e : ev) mn = std::fmin(mn, e);
  if (mn < 1e-10) return ORC_SINGULAR;
  int64_t n = w.r;
  Mat scaled(n, n);
  for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < n; ++j) scaled(i, j) = u(i, j) * (1.0 / std::sqrt(ev[j]));
  out = dot(dot(scaled, false, u, true), false, w, false);
  return ORC_OK;
}

// matrix_exp: math.rs:38-74 (quirk Q7: max-|entry| scaling, <=30 Taylor terms, early exit at 1e-16)
Mat matrix_exp(const Mat& a) {
  int64_t n = a.r;
  double norm = max_abs(a);
  if (norm < 1e-15) return Mat::eye(n);
  int s = (int)std::fmax(std::ceil(std::log2(norm)), 0.0);
  double scale = std::ldexp(1.0, s);
  Mat as(n, n);
  for (int64_t i = 0; i < n * n; ++i) as.p[i] = a.p[i] / scale;
  Mat result = Mat::eye(n), term = Mat::eye(n);
  for (int k = 1; k <= 30; ++k) {
    Mat t2 = dot(term, false, as, false);
    for (int64_t i = 0; i < n * n; ++i) t2.p[i] /= (double)k;
    term = std::move(t2);
    for (int64_t i = 0; i < n * n; ++i) result.p[i] += term.p[i];
    if (max_abs(term) < 1e-16) break;
  }
  for (int i = 0; i < s; ++i) result = dot(result, false, result, false);
  return result;
}

// skew_symmetric: math.rs:91-93
Mat skew(const Mat& a) {
  Mat o(a.r, a.c);
  for (int64_t i = 0; i < a.r; ++i) for (int64_t j = 0; j < a.c; ++j) o(i, j) = (a(i, j) - a(j, i)) / 2.0;
  return o;
}

// ------------------------------------------------------------------------------------------------
// lbfgs.rs
// ------------------------------------------------------------------------------------------------
struct Memory {  // lbfgs.rs:6-16 (only the fields core.rs touches: quirk Q9)
  std::vector<Mat> s, y;
  std::vector<double> r;
  void clear() { s.clear(); y.clear(); r.clear(); }
};
double fdot(const Mat& a, const Mat& b) { double s = 0; for (int64_t i = 0; i < a.size(); ++i) s += a.p[i] * b.p[i]; return s; }

// solve_hessian_system: lbfgs.rs:136-150 (quirk Q15)
Mat solve_hessian_system(const Mat& h, const std::vector<double>& hoff, const Mat& g) {
  int64_t n = h.r;
  Mat out = Mat::zeros(n, n);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j) {
      double det = h(i, j) * h(j, i) - hoff[i] * hoff[j];
      if (std::fabs(det) > 1e-15) out(i, j) = (h(j, i) * g(i, j) - hoff[i] * g(j, i)) / det;
    }
  return out;
}
// regularize_hessian: lbfgs.rs:155-171 (sequential, in place: quirk Q4)
void regularize_hessian(Mat& h, const std::vector<double>& hoff, double lambda_min) {
  int64_t n = h.r;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j)
      if (i != j) {
        double diff = h(i, j) - h(j, i);
        double discr = std::sqrt(diff * diff + 4.0 * hoff[i] * hoff[j]);
        double ev = 0.5 * (h(i, j) + h(j, i) - discr);
        if (ev < lambda_min) h(i, j) += lambda_min - ev;
      }
}
// compute_direction: lbfgs.rs:84-133
Mat compute_direction(const Mat& g, const Mat& h, const std::vector<double>& hoff, const Memory& mem, bool ortho) {
  int64_t nn = g.size();
  Mat q(g);
  size_t L = mem.s.size();
  std::vector<double> al(L);
  for (size_t k = L; k-- > 0;) {
    double a = mem.r[k] * fdot(mem.s[k], q);
    al[k] = a;
    for (int64_t i = 0; i < nn; ++i) q.p[i] = q.p[i] - a * mem.y[k].p[i];
  }
  Mat z;
  if (ortho) {
    Mat z0(g.r, g.c);
    for (int64_t i = 0; i < nn; ++i) z0.p[i] = q.p[i] / h.p[i];
    z = skew(z0);
  } else {
    z = solve_hessian_system(h, hoff, q);
  }
  for (size_t k = 0; k < L; ++k) {
    double beta = mem.r[k] * fdot(mem.y[k], z);
    double c = al[k] - beta;
    for (int64_t i = 0; i < nn; ++i) z.p[i] = z.p[i] + c * mem.s[k].p[i];
  }
  for (int64_t i = 0; i < nn; ++i) z.p[i] = -z.p[i];
  return z;
}

// ------------------------------------------------------------------------------------------------
// core.rs
// ------------------------------------------------------------------------------------------------
struct Counters { int64_t loss_evals = 0, grad_evals = 0, ls_tries = 0, fallbacks = 0; };

enum { LOSS_VALUE = 0, LOSS_SINGULAR = 1, LOSS_ERROR = 2 };
// compute_loss: core.rs:39-85
int compute_loss(const Mat& y, const Mat& w, int kind, double alpha, const std::vector<double>& signs, bool ortho,
                 bool extended, double* out, Counters* cnt) {
  if (cnt) cnt->loss_evals++;
  int64_t n = y.r, t = y.c;
  double tf = (double)t, loss = 0.0;
  if (!ortho) {
    double sg, la;
    if (sln_det(w, &sg, &la) != ORC_OK) return LOSS_ERROR;
    if (sg == 0.0) return LOSS_SINGULAR;
    loss = -la;
  }
  for (int64_t i = 0; i < n; ++i) {
    Mat row(1, t);                                        // row.to_owned()
    memcpy(row.p, y.p + i * t, sizeof(double) * (size_t)t);
    Mat ll(1, t);                                         // density.log_lik(&row)
    log_lik(kind, alpha, row.p, t, ll.p);
    double s = 0; for (int64_t k = 0; k < t; ++k) s += ll.p[k];
    loss += signs[i] * s / tf;
    if (extended && !ortho) {
      double sq = 0; for (int64_t k = 0; k < t; ++k) sq += row.p[k] * row.p[k];
      loss += 0.5 * sq / tf;
    }
  }
  *out = loss;
  return LOSS_VALUE;
}
double loss_to_f64(int st, double v) { return st == LOSS_VALUE ? v : 1e15; }  // core.rs:90-96

struct LineSearch { bool success; Mat y, w; double loss; Mat step; double alpha_used; int tries; };
// line_search: core.rs:99-150 (quirk Q2: on failure alpha has been halved once more than the last try)
LineSearch line_search(const Mat& y, const Mat& w, int kind, double dalpha, const Mat& direction,
                       const std::vector<double>& signs, double current_loss, int64_t ls_tries, bool ortho, bool extended,
                       Counters* cnt) {
  int64_t n = w.r;
  double alpha = 1.0;
  LineSearch R;
  R.y = Mat(y);  // y.clone() (core.rs:113)
  R.w = Mat(w);
  R.loss = current_loss;
  R.tries = 0;
  for (int64_t it = 0; it < ls_tries; ++it) {
    R.tries++;
    if (cnt) cnt->ls_tries++;
    Mat da(n, n);
    for (int64_t i = 0; i < n * n; ++i) da.p[i] = direction.p[i] * alpha;
    Mat transform;
    if (ortho) transform = matrix_exp(da);
    else { transform = Mat::eye(n); for (int64_t i = 0; i < n * n; ++i) transform.p[i] = transform.p[i] + da.p[i]; }
    R.y = dot(transform, false, y, false);
    R.w = dot(transform, false, w, false);
    double v = 0;
    int st = compute_loss(R.y, R.w, kind, dalpha, signs, ortho, extended, &v, cnt);
    R.loss = loss_to_f64(st, v);
    if (R.loss < current_loss) {
      R.success = true; R.step = std::move(da); R.alpha_used = alpha;
      return R;
    }
    alpha /= 2.0;
  }
  R.success = false;
  R.step = Mat(n, n);
  for (int64_t i = 0; i < n * n; ++i) R.step.p[i] = direction.p[i] * alpha;
  R.alpha_used = alpha;
  return R;
}

// The "front half" of one outer iteration, core.rs:215-293: from Y (and C, the previous signs) to the
// projected gradient G, the Hessian approximation h, h_off, signs and the gradient norm.
struct Front {
  Mat g, h;
  std::vector<double> hoff, signs;
  bool sign_change = false;
  double gradient_norm = 0;
  // raw (unsigned, unscaled) moments, for the fused-pass contract SURVEY.md §8(a); not part of the reference
  Mat gr, hr; std::vector<double> sd, sq;
};
Front iteration_front(const Mat& y, int kind, double alpha, bool ortho, bool extended, double lambda_min, const Mat& c,
                      std::vector<double>& old_signs, bool first_iter, bool want_raw, Counters* cnt) {
  if (cnt) cnt->grad_evals++;
  int64_t n = y.r, t = y.c;
  double tf = (double)t;
  Front F;
  Mat psi, psid;
  score_and_der(kind, alpha, y, psi, psid);                               // core.rs:215
  Mat g = dot(psi, false, y, true);                                       // core.rs:218
  if (want_raw) F.gr = Mat(g);
  for (int64_t i = 0; i < n * n; ++i) g.p[i] /= tf;
  Mat ysq(n, t);                                                          // core.rs:221 (also when ortho: Q6)
  for (int64_t i = 0; i < n * t; ++i) ysq.p[i] = y.p[i] * y.p[i];
  if (want_raw) {
    F.sd.assign(n, 0.0); F.sq.assign(n, 0.0);
    for (int64_t i = 0; i < n; ++i) {
      double a = 0, b = 0;
      for (int64_t k = 0; k < t; ++k) { a += psid(i, k); b += ysq(i, k); }
      F.sd[i] = a; F.sq[i] = b;
    }
    F.hr = dot(psid, false, ysq, true);
  }
  F.signs.assign(n, 1.0);
  bool have_signs = false;
  if (extended) {                                                         // core.rs:225-253
    std::vector<double> pm(n);
    for (int64_t i = 0; i < n; ++i) { double s = 0; for (int64_t k = 0; k < t; ++k) s += psid(i, k); pm[i] = s / tf; }
    for (int64_t i = 0; i < n; ++i) F.signs[i] = rust_signum(pm[i] * c(i, i) - g(i, i));
    have_signs = true;
    if (!first_iter)
      for (int64_t i = 0; i < n; ++i) if (F.signs[i] != old_signs[i]) F.sign_change = true;
    old_signs = F.signs;
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t j = 0; j < n; ++j) g(i, j) *= F.signs[i];
      for (int64_t k = 0; k < t; ++k) psid(i, k) *= F.signs[i];
    }
    if (!ortho) {
      for (int64_t i = 0; i < n * n; ++i) g.p[i] = g.p[i] + c.p[i];
      for (int64_t i = 0; i < n * t; ++i) psid.p[i] = psid.p[i] + 1.0;
    }
  }
  (void)have_signs;
  F.hoff.assign(n, 1.0);                                                  // core.rs:256-260
  if (ortho) for (int64_t i = 0; i < n; ++i) F.hoff[i] = g(i, i);
  if (ortho) {                                                            // core.rs:263-272 (Q18)
    std::vector<double> pm(n);
    for (int64_t i = 0; i < n; ++i) { double s = 0; for (int64_t k = 0; k < t; ++k) s += psid(i, k); pm[i] = s / tf; }
    F.h = Mat(n, n);
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < n; ++j) {
        double v = 0.5 * (pm[i] + pm[j] - F.hoff[i] - F.hoff[j]);
        F.h(i, j) = std::fmax(v, lambda_min);
      }
  } else {                                                                // core.rs:274-276
    F.h = dot(psid, false, ysq, true);
    for (int64_t i = 0; i < n * n; ++i) F.h.p[i] /= tf;
    regularize_hessian(F.h, F.hoff, lambda_min);
  }
  if (ortho) g = skew(g);                                                 // core.rs:280-286
  else for (int64_t i = 0; i < n; ++i) g(i, i) -= 1.0;
  F.gradient_norm = max_abs(g);                                           // core.rs:289
  F.g = std::move(g);
  return F;
}

struct TraceRow { double gradient_norm, loss, alpha; int32_t tries, fallback, sign_change, mem_len; };

struct CoreOut {
  Mat y, w;
  bool converged = false;
  double gradient_norm = 1.0;
  int64_t n_iterations = 0;
  std::vector<double> signs;
  bool has_signs = false;
};

// run: core.rs:162-401
int core_run(const Mat& x, int kind, double alpha, bool ortho, bool extended, int64_t m, int64_t max_iter, double tol,
             double lambda_min, int64_t ls_tries, bool verbose, const Mat* covariance, CoreOut& out,
             std::vector<TraceRow>* trace, Counters* cnt) {
  int64_t n = x.r, t = x.c;
  double tf = (double)t;
  Mat w = Mat::eye(n);
  Mat y(x);                                                                // x.clone()
  Memory memory;
  std::vector<double> signs(n, 1.0), old_signs(n, 1.0);
  double current_loss = 0;
  {
    double v = 0;
    int st = compute_loss(y, w, kind, alpha, signs, ortho, extended, &v, cnt);  // core.rs:185-194 (Q1)
    if (st == LOSS_SINGULAR) return ORC_SINGULAR;
    if (st == LOSS_ERROR) return ORC_COMPUTATION;
    current_loss = v;
  }
  double gradient_norm = 1.0;
  bool converged = false;
  Mat c;
  if (extended) {                                                          // core.rs:199-205 (Q5)
    if (covariance) c = Mat(*covariance);
    else { c = dot(y, false, y, true); for (int64_t i = 0; i < n * n; ++i) c.p[i] /= tf; }
  } else c = Mat::eye(n);
  Mat g_old; bool have_g_old = false;
  Mat prev_step; bool have_prev_step = false;
  int64_t n_iter = 0;
  for (int64_t iter = 0; iter < max_iter; ++iter) {
    n_iter = iter;
    Front F = iteration_front(y, kind, alpha, ortho, extended, lambda_min, c, old_signs, iter == 0, false, cnt);
    if (extended) signs = F.signs;
    gradient_norm = F.gradient_norm;
    if (gradient_norm < tol) { converged = true; break; }                  // core.rs:289-293 (Q10)
    if (iter > 0 && have_prev_step && have_g_old) {                        // core.rs:296-313 (Q9)
      Mat step = std::move(prev_step); have_prev_step = false;
      Mat yd(n, n);
      for (int64_t i = 0; i < n * n; ++i) yd.p[i] = F.g.p[i] - g_old.p[i];
      double r = 1.0 / fdot(step, yd);
      if (std::isfinite(r)) {
        memory.s.push_back(std::move(step)); memory.y.push_back(std::move(yd)); memory.r.push_back(r);
        if ((int64_t)memory.s.size() > m) {
          memory.s.erase(memory.s.begin()); memory.y.erase(memory.y.begin()); memory.r.erase(memory.r.begin());
        }
      }
    }
    g_old = Mat(F.g); have_g_old = true;
    if (extended && F.sign_change) {                                       // core.rs:317-331 (Q11)
      double v = 0;
      int st = compute_loss(y, w, kind, alpha, signs, ortho, extended, &v, cnt);
      if (st == LOSS_ERROR) return ORC_COMPUTATION;
      current_loss = (st == LOSS_VALUE) ? v : 1e15;
      memory.clear();
    }
    Mat direction = compute_direction(F.g, F.h, F.hoff, memory, ortho);    // core.rs:334
    int mem_len = (int)memory.s.size();
    LineSearch R = line_search(y, w, kind, alpha, direction, signs, current_loss, ls_tries, ortho, extended, cnt);
    int fallback = 0, tries = R.tries;
    if (!R.success) {                                                      // core.rs:349-367 (Q3)
      fallback = 1;
      if (cnt) cnt->fallbacks++;
      memory.clear();
      Mat neg(n, n);
      for (int64_t i = 0; i < n * n; ++i) neg.p[i] = -F.g.p[i];
      R = line_search(y, w, kind, alpha, neg, signs, current_loss, 10, ortho, extended, cnt);
      tries += R.tries;
    }
    prev_step = std::move(R.step); have_prev_step = true;                  // core.rs:370
    y = std::move(R.y);
    w = std::move(R.w);
    if (extended && covariance) c = dot(dot(w, false, *covariance, false), false, w, true);  // core.rs:375-379
    current_loss = R.loss;
    if (trace) trace->push_back(TraceRow{gradient_norm, current_loss, R.alpha_used, tries, fallback, F.sign_change ? 1 : 0, mem_len});
    if (verbose) {
      printf("iteration %lld, gradient norm = %s, loss = %s\n", (long long)(iter + 1), rust_e4(gradient_norm).c_str(),
             rust_e4(current_loss).c_str());
      fflush(stdout);
    }
  }
  out.y = std::move(y); out.w = std::move(w);
  out.converged = converged; out.gradient_norm = gradient_norm;
  out.n_iterations = n_iter + 1;
  out.has_signs = extended; out.signs = signs;
  return ORC_OK;
}

// ------------------------------------------------------------------------------------------------
// whitening.rs
// ------------------------------------------------------------------------------------------------
// center: whitening.rs:24-35
void center(const Mat& x, Mat& centered, std::vector<double>& mean) {
  int64_t n = x.r, t = x.c;
  mean.assign(n, 0.0);
  for (int64_t i = 0; i < n; ++i) { double s = 0; for (int64_t k = 0; k < t; ++k) s += x(i, k); mean[i] = s / (double)t; }
  centered = Mat(x);
  for (int64_t i = 0; i < n; ++i) for (int64_t k = 0; k < t; ++k) centered(i, k) -= mean[i];
}
// whiten: whitening.rs:48-116 (thin SVD, U only; sign rule Q16)
int whiten(const Mat& x, int64_t n_components, Mat& data, Mat& kmat) {
  int64_t nf = x.r, t = x.c;
  if (n_components > nf) return ORC_INVALID_DIMENSIONS;
  // Row-major x (nf x t) is the column-major matrix x^T (t x nf).  x^T = U' S V'^T  =>  x = V' S U'^T:
  // the left singular vectors of x are the rows of LAPACK's VT for x^T.  (ndarray-linalg does the same swap.)
  Mat a(x);
  int mm = (int)t, nn = (int)nf, lda = mm, ldu = 1, ldvt = nn, info = 0, lwork = -1;
  int kmin = mm < nn ? mm : nn;
  std::vector<double> s((size_t)kmin), vt((size_t)nn * nn);
  double wq = 0, udummy = 0;
  scipy_dgesvd_("N", "A", &mm, &nn, a.p, &lda, s.data(), &udummy, &ldu, vt.data(), &ldvt, &wq, &lwork, &info, 1, 1);
  lwork = (int)wq;
  std::vector<double> work((size_t)(lwork > 1 ? lwork : 1));
  scipy_dgesvd_("N", "A", &mm, &nn, a.p, &lda, s.data(), &udummy, &ldu, vt.data(), &ldvt, work.data(), &lwork, &info, 1, 1);
  if (info != 0) return ORC_COMPUTATION;
  double mn = std::numeric_limits<double>::infinity();
  for (int64_t i = 0; i < n_components && i < kmin; ++i) mn = std::fmin(mn, s[i]);
  if (mn < 1e-10) return ORC_SINGULAR;
  double scale = std::sqrt((double)t);
  kmat = Mat::zeros(n_components, nf);
  // u[j][i] (component j of singular vector i) = VT'(i, j) = vt[i + j*nn] (column-major)
  for (int64_t i = 0; i < n_components; ++i)
    for (int64_t j = 0; j < nf; ++j) kmat(i, j) = vt[(size_t)i + (size_t)j * nn] / s[i] * scale;
  for (int64_t i = 0; i < n_components; ++i) {                            // whitening.rs:93-107
    int64_t best = 0;
    // Iterator::max_by returns the LAST maximum on ties
    for (int64_t j = 0; j < nf; ++j) if (std::fabs(kmat(i, j)) >= std::fabs(kmat(i, best))) best = j;
    if (kmat(i, best) < 0.0) for (int64_t j = 0; j < nf; ++j) kmat(i, j) = -kmat(i, j);
  }
  data = dot(kmat, false, x, false);
  return ORC_OK;
}

// ------------------------------------------------------------------------------------------------
// jade.rs
// ------------------------------------------------------------------------------------------------
// compute_cumulant_matrices: jade.rs:78-131 (materialises x_i x_j; N(N+1)/2 matrices)
std::vector<Mat> cumulant_matrices(const Mat& x) {
  int64_t n = x.r, t = x.c;
  double tf = (double)t;
  std::vector<Mat> out;
  std::vector<double> xx((size_t)(n * n * t));
  for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < n; ++j) for (int64_t s = 0; s < t; ++s) xx[(size_t)((i * n + j) * t + s)] = x(i, s) * x(j, s);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = i; j < n; ++j) {
      Mat q = Mat::zeros(n, n);
      const double* pij = &xx[(size_t)((i * n + j) * t)];
      for (int64_t k = 0; k < n; ++k)
        for (int64_t l = 0; l < n; ++l) {
          const double* pkl = &xx[(size_t)((k * n + l) * t)];
          double e = 0.0;
          for (int64_t s = 0; s < t; ++s) e += pij[s] * pkl[s];
          e /= tf;
          double d1 = (i == j && k == l) ? 1.0 : 0.0, d2 = (i == k && j == l) ? 1.0 : 0.0, d3 = (i == l && j == k) ? 1.0 : 0.0;
          q(k, l) = e - d1 - d2 - d3;
        }
      Mat qs(n, n);
      for (int64_t k = 0; k < n; ++k) for (int64_t l = 0; l < n; ++l) qs(k, l) = (q(k, l) + q(l, k)) / 2.0;
      out.push_back(std::move(qs));
    }
  return out;
}
// compute_givens_rotation: jade.rs:137-185 ; apply_givens_rotation: jade.rs:188-197 (quirk Q17)
void givens(const std::vector<Mat>& ms, const Mat& v, int64_t p, int64_t q, double* c, double* s, double* theta) {
  double g00 = 0, g01 = 0, g11 = 0;
  int64_t n = v.r;
  const int64_t idx[2] = {p, q};
  for (const Mat& m : ms) {
    double b[2][2] = {{0, 0}, {0, 0}};
    for (int bi = 0; bi < 2; ++bi)
      for (int bj = 0; bj < 2; ++bj)
        for (int64_t k = 0; k < n; ++k)
          for (int64_t l = 0; l < n; ++l) b[bi][bj] += v(k, idx[bi]) * m(k, l) * v(l, idx[bj]);
    double hpq = b[0][1] + b[1][0], hd = b[0][0] - b[1][1];
    g00 += hpq * hpq; g01 += hpq * hd; g11 += hd * hd;
  }
  double diff = g11 - g00, ang;
  if (std::fabs(g01) < 1e-15 && std::fabs(diff) < 1e-15) ang = 0.0;
  else ang = 0.25 * std::atan2(2.0 * g01, diff);
  *c = std::cos(ang); *s = std::sin(ang); *theta = ang;
}
// jade: jade.rs:22-72 (returns sym_decorrelation(V), not V^T)
int jade(const Mat& x, int64_t max_iter, double tol, bool verbose, Mat& w, int64_t* sweeps_done) {
  int64_t n = x.r;
  if (sweeps_done) *sweeps_done = 0;
  if (n < 2) { w = Mat::eye(n); return ORC_OK; }
  std::vector<Mat> cum = cumulant_matrices(x);
  if (verbose) printf("JADE: %zu cumulant matrices computed\n", cum.size());
  Mat v = Mat::eye(n);
  for (int64_t it = 0; it < max_iter; ++it) {
    double max_theta = 0.0;
    for (int64_t p = 0; p < n; ++p)
      for (int64_t q = p + 1; q < n; ++q) {
        double c, s, th;
        givens(cum, v, p, q, &c, &s, &th);
        max_theta = std::fmax(max_theta, std::fabs(th));
        for (int64_t i = 0; i < n; ++i) {
          double vp = v(i, p), vq = v(i, q);
          v(i, p) = c * vp - s * vq;
          v(i, q) = s * vp + c * vq;
        }
      }
    if (sweeps_done) *sweeps_done = it + 1;
    if (verbose && (it + 1) % 10 == 0) printf("JADE iteration %lld: max angle = %s\n", (long long)(it + 1), rust_e4(max_theta).c_str());
    if (max_theta < tol) {
      if (verbose) printf("JADE converged after %lld iterations\n", (long long)(it + 1));
      break;
    }
  }
  return sym_decorrelation(v, w);
}

// ------------------------------------------------------------------------------------------------
// the build's own documented generator for the random w_init path (NOT rand's ChaCha12/Ziggurat):
// splitmix64 stream; u = (next >> 11 + 0.5) * 2^-53 in (0,1); Box-Muller, both outputs used in order.
// The CUDA library's host code implements the same specification independently.
// ------------------------------------------------------------------------------------------------
struct orc_rng {
  uint64_t s; bool have = false; double spare = 0;
  explicit orc_rng(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double normal() {
    if (have) { have = false; return spare; }
    double u1 = uniform(), u2 = uniform();
    double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586476925286766559 * u2;
    spare = r * std::sin(a); have = true;
    return r * std::cos(a);
  }
};

// ica_par: solver.rs:218-249
int ica_par(const Mat& x, int kind, double alpha, int64_t max_iter, const Mat& w_init, bool verbose, Mat& wout) {
  Mat w;
  int st = sym_decorrelation(w_init, w);
  if (st != ORC_OK) return st;
  double p = (double)x.c;
  int64_t n = w.r, t = x.c;
  for (int64_t it = 0; it < max_iter; ++it) {
    Mat wx = dot(w, false, x, false);
    Mat gw, gpw;
    score_and_der(kind, alpha, wx, gw, gpw);
    std::vector<double> gm(n);
    for (int64_t i = 0; i < n; ++i) { double s = 0; for (int64_t k = 0; k < t; ++k) s += gpw(i, k); gm[i] = s / (double)t; }
    Mat c = dot(gw, false, x, true);
    for (int64_t i = 0; i < c.size(); ++i) c.p[i] /= p;
    for (int64_t i = 0; i < w.r; ++i) for (int64_t j = 0; j < w.c; ++j) c(i, j) -= gm[i] * w(i, j);
    st = sym_decorrelation(c, w);
    if (st != ORC_OK) return st;
  }
  if (verbose) printf("FastICA pre-iterations complete.\n");
  wout = std::move(w);
  return ORC_OK;
}

}  // namespace

// =====================================================================================================
// C interface (ctypes).  All matrices row-major, f64.
// =====================================================================================================
extern "C" {

struct orc_config {       // mirrors PicardConfig, config.rs:11-62 (same field meaning as picard_config_t)
  int32_t density_kind; double alpha;
  int64_t n_components;   // -1 = None
  int32_t ortho, extended /* -1 = None */, whiten, centering;
  int64_t max_iter; double tol; int64_t m, ls_tries; double lambda_min;
  const double* w_init;   // nc x nc or NULL
  int64_t fastica_it, jade_it;  // -1 = None
  int32_t has_seed; uint64_t seed; int32_t verbose;
};
struct orc_result {       // mirrors PicardResult, result.rs:7-33; buffers malloc'd here, freed by orc_result_free
  int64_t n_components, n_features, n_samples;
  double* whitening;      // nc x nf or NULL
  double* unmixing;       // nc x nc
  double* sources;        // nc x T
  double* mean;           // nf or NULL
  int64_t n_iterations; int32_t converged; double gradient_norm;
  double* signs;          // nc or NULL
  // extras for parity/measurement (not in the reference's struct)
  double* w_init_used;    // nc x nc, the w_init actually applied (after warm start)
  int64_t loss_evals, grad_evals, ls_tries_total, fallbacks;
  double* trace; int64_t trace_rows;  // rows of 7 doubles: gradient_norm, loss, alpha, tries, fallback, sign_change, mem_len
  double core_seconds;
};

void orc_set_threads(int n) { scipy_openblas_set_num_threads(n); }
int orc_get_threads(void) { return scipy_openblas_get_num_threads(); }

void orc_log_lik(int kind, double alpha, const double* y, int64_t n, double* out) { log_lik(kind, alpha, y, n, out); }
void orc_score_and_der(int kind, double alpha, const double* y, int64_t r, int64_t c, double* psi, double* psid) {
  Mat Y = Mat::from(y, r, c), a, b;
  score_and_der(kind, alpha, Y, a, b);
  memcpy(psi, a.p, sizeof(double) * (size_t)(r * c));
  memcpy(psid, b.p, sizeof(double) * (size_t)(r * c));
}
int orc_sln_det(const double* m, int64_t n, double* sign, double* logabs) { return sln_det(Mat::from(m, n, n), sign, logabs); }
int orc_sym_decorrelation(const double* w, int64_t n, double* out) {
  Mat o; int st = sym_decorrelation(Mat::from(w, n, n), o);
  if (st == ORC_OK) memcpy(out, o.p, sizeof(double) * (size_t)(n * n));
  return st;
}
void orc_matrix_exp(const double* a, int64_t n, double* out) { Mat o = matrix_exp(Mat::from(a, n, n)); memcpy(out, o.p, sizeof(double) * (size_t)(n * n)); }
void orc_skew(const double* a, int64_t n, double* out) { Mat o = skew(Mat::from(a, n, n)); memcpy(out, o.p, sizeof(double) * (size_t)(n * n)); }
void orc_regularize_hessian(double* h, const double* hoff, int64_t n, double lambda_min) {
  Mat H = Mat::from(h, n, n); std::vector<double> ho(hoff, hoff + n);
  regularize_hessian(H, ho, lambda_min);
  memcpy(h, H.p, sizeof(double) * (size_t)(n * n));
}
// compute_direction with an explicit memory (s_list, y_list: L x n x n; r_list: L), oldest first
void orc_compute_direction(const double* g, const double* h, const double* hoff, int64_t n, const double* s_list,
                           const double* y_list, const double* r_list, int64_t L, int ortho, double* out) {
  Memory mem;
  for (int64_t k = 0; k < L; ++k) {
    mem.s.push_back(Mat::from(s_list + k * n * n, n, n)); mem.y.push_back(Mat::from(y_list + k * n * n, n, n)); mem.r.push_back(r_list[k]);
  }
  std::vector<double> ho(hoff, hoff + n);
  Mat d = compute_direction(Mat::from(g, n, n), Mat::from(h, n, n), ho, mem, ortho != 0);
  memcpy(out, d.p, sizeof(double) * (size_t)(n * n));
}
// loss at Y = W X with the given signs; returns status LOSS_*; *out = value (1e15 on singular, as loss_to_f64)
int orc_compute_loss(const double* y, const double* w, int64_t n, int64_t t, int kind, double alpha, const double* signs,
                     int ortho, int extended, double* out) {
  std::vector<double> sg(signs, signs + n);
  double v = 0;
  int st = compute_loss(Mat::from(y, n, t), Mat::from(w, n, n), kind, alpha, sg, ortho != 0, extended != 0, &v, nullptr);
  *out = loss_to_f64(st, v);
  return st;
}

// One evaluation point: Y = W X (W may be NULL = identity), then the literal front half of an iteration
// (core.rs:215-293) and the loss (core.rs:39-85) with `loss_signs` (NULL = the signs just estimated, or
// ones when not extended).  Also returns the raw moments of the fused-pass contract.
// c: covariance-like matrix C (N x N) used by the extended sign rule; NULL = identity.
// old_signs: NULL = first iteration (sign_change forced false).
int orc_eval_point(const double* x, int64_t n, int64_t t, const double* w, int kind, double alpha, int ortho, int extended,
                   double lambda_min, const double* c, const double* old_signs, const double* loss_signs,
                   /* raw */ double* gr, double* sd, double* hr, double* sq, double* lrow,
                   /* processed */ double* g, double* h, double* hoff, double* signs, int32_t* sign_change,
                   double* gradient_norm, double* loss, int32_t* loss_status) {
  Mat X = Mat::from(x, n, t);
  Mat W = w ? Mat::from(w, n, n) : Mat::eye(n);
  Mat Y = w ? dot(W, false, X, false) : Mat(X);
  Mat C = c ? Mat::from(c, n, n) : Mat::eye(n);
  std::vector<double> os(n, 1.0);
  if (old_signs) os.assign(old_signs, old_signs + n);
  Front F = iteration_front(Y, kind, alpha, ortho != 0, extended != 0, lambda_min, C, os, old_signs == nullptr, true, nullptr);
  if (gr) memcpy(gr, F.gr.p, sizeof(double) * (size_t)(n * n));
  if (hr) memcpy(hr, F.hr.p, sizeof(double) * (size_t)(n * n));
  if (sd) memcpy(sd, F.sd.data(), sizeof(double) * (size_t)n);
  if (sq) memcpy(sq, F.sq.data(), sizeof(double) * (size_t)n);
  if (lrow) {
    std::vector<double> tmp((size_t)t);
    for (int64_t i = 0; i < n; ++i) { log_lik(kind, alpha, Y.p + i * t, t, tmp.data()); double s = 0; for (int64_t k = 0; k < t; ++k) s += tmp[k]; lrow[i] = s; }
  }
  if (g) memcpy(g, F.g.p, sizeof(double) * (size_t)(n * n));
  if (h) memcpy(h, F.h.p, sizeof(double) * (size_t)(n * n));
  if (hoff) memcpy(hoff, F.hoff.data(), sizeof(double) * (size_t)n);
  if (signs) memcpy(signs, F.signs.data(), sizeof(double) * (size_t)n);
  if (sign_change) *sign_change = F.sign_change ? 1 : 0;
  if (gradient_norm) *gradient_norm = F.gradient_norm;
  if (loss) {
    std::vector<double> ls = loss_signs ? std::vector<double>(loss_signs, loss_signs + n) : F.signs;
    double v = 0;
    int st = compute_loss(Y, W, kind, alpha, ls, ortho != 0, extended != 0, &v, nullptr);
    *loss = loss_to_f64(st, v);
    if (loss_status) *loss_status = st;
  }
  return ORC_OK;
}

void orc_center(const double* x, int64_t n, int64_t t, double* centered, double* mean) {
  Mat c; std::vector<double> mu;
  center(Mat::from(x, n, t), c, mu);
  memcpy(centered, c.p, sizeof(double) * (size_t)(n * t));
  memcpy(mean, mu.data(), sizeof(double) * (size_t)n);
}
int orc_whiten(const double* x, int64_t nf, int64_t t, int64_t nc, double* data, double* k) {
  Mat d, K;
  int st = whiten(Mat::from(x, nf, t), nc, d, K);
  if (st != ORC_OK) return st;
  memcpy(data, d.p, sizeof(double) * (size_t)(nc * t));
  memcpy(k, K.p, sizeof(double) * (size_t)(nc * nf));
  return ORC_OK;
}
// cumulant matrices, [n(n+1)/2] x n x n, order (i, j>=i)
void orc_cumulants(const double* x, int64_t n, int64_t t, double* out) {
  std::vector<Mat> c = cumulant_matrices(Mat::from(x, n, t));
  for (size_t k = 0; k < c.size(); ++k) memcpy(out + k * n * n, c[k].p, sizeof(double) * (size_t)(n * n));
}
int orc_jade(const double* x, int64_t n, int64_t t, int64_t max_iter, double tol, int verbose, double* w, int64_t* sweeps) {
  Mat W;
  int st = jade(Mat::from(x, n, t), max_iter, tol, verbose != 0, W, sweeps);
  if (st == ORC_OK) memcpy(w, W.p, sizeof(double) * (size_t)(n * n));
  return st;
}
// amari_distance: utils.rs:82-103
double orc_amari(const double* w, const double* a, int64_t n) {
  Mat P = dot(Mat::from(w, n, n), false, Mat::from(a, n, n), false);
  auto srow = [&](bool transpose) {
    double sum = 0;
    for (int64_t i = 0; i < n; ++i) {
      double rs = 0, rm = 0;
      for (int64_t j = 0; j < n; ++j) { double v = std::fabs(transpose ? P(j, i) : P(i, j)); v = v * v; rs += v; rm = std::fmax(rm, v); }
      if (rm > 1e-15) sum += rs / rm - 1.0;
    }
    return sum;
  };
  return (srow(false) + srow(true)) / (2.0 * (double)n);
}
// N(0,1) matrix from the build's generator, row-major fill order (solver.rs:113-119 shape)
void orc_randn(uint64_t seed, int64_t n, double* out) { orc_rng g(seed); for (int64_t i = 0; i < n; ++i) out[i] = g.normal(); }

// config.validate(): config.rs:104-142
int orc_validate(const orc_config* c, char* err, size_t errlen) {
  auto fail = [&](const char* param, const char* msg) { if (err && errlen) snprintf(err, errlen, "Invalid configuration for '%s': %s", param, msg); return ORC_INVALID_CONFIG; };
  if (c->max_iter <= 0) return fail("max_iter", "must be greater than 0");
  if (!(c->tol > 0.0)) return fail("tol", "must be positive");
  if (!(c->lambda_min > 0.0)) return fail("lambda_min", "must be positive");
  if (c->m <= 0) return fail("m", "L-BFGS memory size must be at least 1");
  if (c->fastica_it >= 0 && c->jade_it >= 0) return fail("jade_it", "cannot use both fastica_it and jade_it; choose one warm start method");
  return ORC_OK;
}

void orc_result_free(orc_result* r) {
  if (!r) return;
  free(r->whitening); free(r->unmixing); free(r->sources); free(r->mean); free(r->signs); free(r->w_init_used); free(r->trace);
  memset(r, 0, sizeof *r);
}

static double* dup_buf(const double* p, size_t n) { double* o = (double*)malloc(sizeof(double) * (n ? n : 1)); memcpy(o, p, sizeof(double) * n); return o; }

// Picard::fit_with_config: solver.rs:45-189
int orc_fit(const double* x, int64_t n, int64_t p, const orc_config* cfg, orc_result* out, char* err, size_t errlen) {
  memset(out, 0, sizeof *out);
  int st = orc_validate(cfg, err, errlen);
  if (st != ORC_OK) return st;
  auto fail = [&](int code, const std::string& msg) { if (err && errlen) snprintf(err, errlen, "%s", msg.c_str()); return code; };
  if (n <= 0 || p <= 0) return fail(ORC_INVALID_DIMENSIONS, "Invalid dimensions: Input matrix cannot be empty");
  int64_t mn = n < p ? n : p;
  int64_t ncomp = cfg->n_components >= 0 ? cfg->n_components : mn;
  if (ncomp > mn) ncomp = mn;
  bool ortho = cfg->ortho != 0;
  bool extended = cfg->extended < 0 ? ortho : (cfg->extended != 0);
  if (cfg->density_kind != DENS_TANH && extended && !ortho)
    fprintf(stderr, "Warning: Using a density other than tanh with extended=true and ortho=false may result in incorrect estimation or numerical overflow\n");
  Mat X = Mat::from(x, n, p);
  Mat x1; std::vector<double> mean; bool has_mean = false;
  if (cfg->centering) { center(X, x1, mean); has_mean = true; } else x1 = Mat(X);
  Mat K; bool has_k = false;
  if (cfg->whiten) {
    Mat d;
    st = whiten(x1, ncomp, d, K);
    if (st == ORC_INVALID_DIMENSIONS) return fail(st, "Invalid dimensions: n_components cannot exceed n_features");
    if (st == ORC_SINGULAR) return fail(st, "Singular matrix encountered during computation");
    if (st != ORC_OK) return fail(st, "Computation error: SVD failed");
    x1 = std::move(d); has_k = true;
  }
  int64_t nc = x1.r;
  Mat w_init;
  if (cfg->w_init) w_init = Mat::from(cfg->w_init, nc, nc);   // shape is the caller's contract at this ABI
  else {
    uint64_t seed = cfg->has_seed ? cfg->seed : (uint64_t)std::rand() * 2654435761ull;
    Mat g(nc, nc); orc_randn(seed, nc * nc, g.p);
    st = sym_decorrelation(g, w_init);
    if (st != ORC_OK) return fail(st, st == ORC_SINGULAR ? "Singular matrix encountered during computation" : "Computation error: Eigendecomposition failed in symmetric decorrelation");
  }
  if (cfg->jade_it >= 0) {
    if (cfg->verbose) printf("Running %lld iterations of JADE...\n", (long long)cfg->jade_it);
    Mat wj; st = jade(x1, cfg->jade_it, 1e-6, cfg->verbose != 0, wj, nullptr);
    if (st != ORC_OK) return fail(st, "JADE failed");
    w_init = std::move(wj);
  } else if (cfg->fastica_it >= 0) {
    if (cfg->verbose) printf("Running %lld iterations of FastICA...\n", (long long)cfg->fastica_it);
    Mat wf; st = ica_par(x1, cfg->density_kind, cfg->alpha, cfg->fastica_it, w_init, cfg->verbose != 0, wf);
    if (st != ORC_OK) return fail(st, "FastICA failed");
    w_init = std::move(wf);
  }
  x1 = dot(w_init, false, x1, false);                                     // solver.rs:140
  Mat cov; const Mat* covp = nullptr;
  if (extended && cfg->whiten) { cov = Mat::eye(nc); covp = &cov; }       // solver.rs:143-147
  if (cfg->verbose) printf("Running Picard...\n");
  CoreOut co; std::vector<TraceRow> trace; Counters cnt;
  struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
  st = core_run(x1, cfg->density_kind, cfg->alpha, ortho, extended, cfg->m, cfg->max_iter, cfg->tol, cfg->lambda_min,
                cfg->ls_tries, cfg->verbose != 0, covp, co, &trace, &cnt);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (st == ORC_SINGULAR) return fail(st, "Singular matrix encountered during computation");
  if (st != ORC_OK) return fail(st, "Computation error: LU decomposition failed in determinant computation");
  Mat wfull = dot(co.w, false, w_init, false);                            // solver.rs:169
  if (!co.converged && cfg->verbose)
    fprintf(stderr, "Warning: PICARD did not converge. Final gradient norm: %s, tolerance: %s\n", rust_e4(co.gradient_norm).c_str(), rust_e4(cfg->tol).c_str());
  out->n_components = nc; out->n_features = n; out->n_samples = p;
  out->whitening = has_k ? dup_buf(K.p, (size_t)(nc * n)) : nullptr;
  out->unmixing = dup_buf(wfull.p, (size_t)(nc * nc));
  out->sources = co.y.p; co.y.p = nullptr;
  out->mean = has_mean ? dup_buf(mean.data(), (size_t)n) : nullptr;
  out->n_iterations = co.n_iterations; out->converged = co.converged ? 1 : 0; out->gradient_norm = co.gradient_norm;
  out->signs = co.has_signs ? dup_buf(co.signs.data(), (size_t)nc) : nullptr;
  out->w_init_used = dup_buf(w_init.p, (size_t)(nc * nc));
  out->loss_evals = cnt.loss_evals; out->grad_evals = cnt.grad_evals; out->ls_tries_total = cnt.ls_tries; out->fallbacks = cnt.fallbacks;
  out->trace_rows = (int64_t)trace.size();
  out->trace = (double*)malloc(sizeof(double) * 7 * (trace.size() ? trace.size() : 1));
  for (size_t i = 0; i < trace.size(); ++i) {
    double* r = out->trace + 7 * i;
    r[0] = trace[i].gradient_norm; r[1] = trace[i].loss; r[2] = trace[i].alpha; r[3] = trace[i].tries; r[4] = trace[i].fallback; r[5] = trace[i].sign_change; r[6] = trace[i].mem_len;
  }
  out->core_seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return ORC_OK;
}

// Picard::transform: solver.rs:199-214 with full_unmixing (result.rs:39-44)
void orc_transform(const double* x, int64_t nf, int64_t t, const double* mean, const double* whitening, const double* unmixing,
                   int64_t nc, double* out) {
  Mat X = Mat::from(x, nf, t);
  if (mean) for (int64_t i = 0; i < nf; ++i) for (int64_t k = 0; k < t; ++k) X(i, k) -= mean[i];
  Mat W = whitening ? dot(Mat::from(unmixing, nc, nc), false, Mat::from(whitening, nc, nf), false) : Mat::from(unmixing, nc, nc);
  Mat Y = dot(W, false, X, false);
  memcpy(out, Y.p, sizeof(double) * (size_t)(Y.r * Y.c));
}

// core::run alone on already-preprocessed data (what the GPU core loop is compared with pass by pass)
int orc_core_run(const double* x, int64_t n, int64_t t, int kind, double alpha, int ortho, int extended, int64_t m,
                 int64_t max_iter, double tol, double lambda_min, int64_t ls_tries, int verbose, const double* cov,
                 double* y_out, double* w_out, int32_t* converged, double* gradient_norm, int64_t* n_iterations, double* signs,
                 double* trace /* max_iter x 7 or NULL */, int64_t* trace_rows, int64_t* counters /* 4 or NULL */, double* seconds) {
  Mat C; const Mat* cp = nullptr;
  if (cov) { C = Mat::from(cov, n, n); cp = &C; }
  CoreOut co; std::vector<TraceRow> tr; Counters cnt;
  struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
  int st = core_run(Mat::from(x, n, t), kind, alpha, ortho != 0, extended != 0, m, max_iter, tol, lambda_min, ls_tries, verbose != 0, cp, co, &tr, &cnt);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (st != ORC_OK) return st;
  if (y_out) memcpy(y_out, co.y.p, sizeof(double) * (size_t)(n * t));
  if (w_out) memcpy(w_out, co.w.p, sizeof(double) * (size_t)(n * n));
  if (converged) *converged = co.converged;
  if (gradient_norm) *gradient_norm = co.gradient_norm;
  if (n_iterations) *n_iterations = co.n_iterations;
  if (signs) memcpy(signs, co.signs.data(), sizeof(double) * (size_t)n);
  if (trace) for (size_t i = 0; i < tr.size(); ++i) {
    double* r = trace + 7 * i;
    r[0] = tr[i].gradient_norm; r[1] = tr[i].loss; r[2] = tr[i].alpha; r[3] = tr[i].tries; r[4] = tr[i].fallback; r[5] = tr[i].sign_change; r[6] = tr[i].mem_len;
  }
  if (trace_rows) *trace_rows = (int64_t)tr.size();
  if (counters) { counters[0] = cnt.loss_evals; counters[1] = cnt.grad_evals; counters[2] = cnt.ls_tries; counters[3] = cnt.fallbacks; }
  return ORC_OK;
}

}  // extern "C"
