// The fit pipeline: Picard::fit_with_config (solver.rs:45-189), Picard::transform (solver.rs:199-214),
// centering + whitening (whitening.rs:24-116) with every N x T step on the device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <random>
#include <mutex>
#include <thread>

#include "engine.cuh"
#include "jade.cuh"

namespace picard {

void config_default(picard_config_t* c) {  // config.rs:64-85
  memset(c, 0, sizeof *c);
  c->density_kind = PICARD_DENSITY_TANH; c->alpha = 1.0;
  c->n_components = -1; c->ortho = 1; c->extended = -1; c->whiten = 1; c->centering = 1;
  c->max_iter = 500; c->tol = 1e-7; c->m = 7; c->ls_tries = 10; c->lambda_min = 0.01;
  c->w_init = nullptr; c->w_init_rows = 0; c->w_init_cols = 0; c->fastica_it = -1; c->jade_it = -1;
  c->has_seed = 0; c->seed = 0; c->verbose = 0; c->device = -1; c->comm = nullptr; c->flags = 0;
}

void config_validate(const picard_config_t& c) {  // config.rs:104-142, same order, same messages
  auto fail = [](const char* param, const char* msg) {
    throw Error(PICARD_INVALID_CONFIG, std::string("Invalid configuration for '") + param + "': " + msg);
  };
  if (c.max_iter <= 0) fail("max_iter", "must be greater than 0");
  if (!(c.tol > 0.0)) fail("tol", "must be positive");
  if (!(c.lambda_min > 0.0)) fail("lambda_min", "must be positive");
  if (c.m <= 0) fail("m", "L-BFGS memory size must be at least 1");
  if (c.fastica_it >= 0 && c.jade_it >= 0) fail("jade_it", "cannot use both fastica_it and jade_it; choose one warm start method");
  if (c.density_kind < 0 || c.density_kind > 2) fail("density", "unknown density kind");
  // not checked by the reference (density.rs:31-34,72-75 accept any alpha); alpha <= 0 makes exp(-2 alpha |y|) overflow there
  // and is rejected loudly here because the device kernels rely on exp arguments <= 0
  if (c.density_kind != PICARD_DENSITY_CUBE && !(c.alpha > 0.0)) fail("density", "alpha must be positive");
}

// splitmix64 stream; u = ((next >> 11) + 0.5) 2^-53; Box-Muller, both outputs used in order.  This is the
// build's own generator (the reference's rand 0.9 ChaCha12/Ziggurat stream is not reproducible here):
// same distribution, different stream.  Parity runs pass w_init explicitly.
void randn_fill(uint64_t seed, double* out, size_t count) {
  uint64_t s = seed;
  auto next = [&]() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  auto uni = [&]() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); };
  size_t i = 0;
  while (i < count) {
    const double u1 = uni(), u2 = uni();
    const double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586476925286766559 * u2;
    out[i++] = r * std::cos(a);
    if (i < count) out[i++] = r * std::sin(a);
  }
}

static void host_matmul(const double* a, const double* b, double* c, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += a[(size_t)i * k + l] * b[(size_t)l * n + j];
      c[(size_t)i * n + j] = s;
    }
}

int apply_device(const double* a_host, const double* mean_host, int n_out, int n_in, const double* d_in, int64_t ld_in,
                 double* d_out, int64_t ld_out, int64_t t_local, int sm_count, cudaStream_t st) {
  if ((ld_out & 1) || (reinterpret_cast<uintptr_t>(d_out) & 15))
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: device output must be 16-byte aligned with an even row stride");
  DevBuf<double> dA((size_t)n_out * n_in), dB((size_t)n_out);
  std::vector<double> bias((size_t)n_out, 0.0);
  if (mean_host)
    for (int i = 0; i < n_out; ++i) { double s = 0; for (int k = 0; k < n_in; ++k) s += a_host[(size_t)i * n_in + k] * mean_host[k]; bias[i] = s; }
  PICARD_CUDA(cudaMemcpyAsync(dA.p, a_host, sizeof(double) * n_out * n_in, cudaMemcpyHostToDevice, st));
  PICARD_CUDA(cudaMemcpyAsync(dB.p, bias.data(), sizeof(double) * n_out, cudaMemcpyHostToDevice, st));
  PassLaunch L;
  L.d_x = d_in; L.ldx = ld_in; L.t_local = t_local; L.n_in = n_in; L.n_out = n_out; L.d_w = dA.p; L.ldw = n_in;
  L.d_bias = mean_host ? dB.p : nullptr; L.dens = DENS_LINEAR; L.alpha = 1.0; L.mode = PASS_APPLY; L.want_h = false;
  L.d_partial = nullptr; L.d_mom = nullptr; L.d_out = d_out; L.ld_out = ld_out; L.sm_count = sm_count; L.stream = st;
  int launches = launch_pass(L);
  PICARD_CUDA(cudaStreamSynchronize(st));  // dA / dB go out of scope
  return launches;
}

// ---- host-side N x N helpers of the whitening refinement (rare path: ill-conditioned data) ------------------------------
namespace {
// cyclic Jacobi eigendecomposition of a symmetric matrix (row-major, destroyed): a = V diag(w) V^T, V in columns
void host_jacobi_eigh(std::vector<double>& a, int n, std::vector<double>& w, std::vector<double>& v) {
  v.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) { diag += a[(size_t)i * n + i] * a[(size_t)i * n + i]; for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j]; }
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - sn * akq; a[(size_t)k * n + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - sn * aqk; a[(size_t)q * n + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * vkp - sn * vkq; v[(size_t)k * n + q] = sn * vkp + c * vkq;
        }
      }
  }
  w.resize((size_t)n);
  for (int i = 0; i < n; ++i) w[i] = a[(size_t)i * n + i];
}
// one-sided (Hestenes) Jacobi: rotates the COLUMNS of b (n x n, row-major) until they are mutually orthogonal, accumulating the
// rotations in j (b_in j = b_out): singular values = column norms of b_out, right singular vectors of b_in = columns of j.
// For b = (well-conditioned) x diag(graded) the small singular values come out with high RELATIVE accuracy (Demmel-Veselic).
void host_one_sided_jacobi(std::vector<double>& b, int n, std::vector<double>& j) {
  j.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) j[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double app = 0, aqq = 0, apq = 0;
        for (int k = 0; k < n; ++k) { const double x = b[(size_t)k * n + p], y = b[(size_t)k * n + q]; app += x * x; aqq += y * y; apq += x * y; }
        if (std::fabs(apq) <= 1e-15 * std::sqrt(app * aqq) || apq == 0.0) continue;
        rotated = true;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {
          const double x = b[(size_t)k * n + p], y = b[(size_t)k * n + q];
          b[(size_t)k * n + p] = c * x - sn * y; b[(size_t)k * n + q] = sn * x + c * y;
          const double u = j[(size_t)k * n + p], w2 = j[(size_t)k * n + q];
          j[(size_t)k * n + p] = c * u - sn * w2; j[(size_t)k * n + q] = sn * u + c * w2;
        }
      }
    if (!rotated) break;
  }
}
}  // namespace

void center_whiten_device(const double* d_x, int nf, int64_t t_local, int64_t ldx, int nc, bool centering, bool whiten,
                          picard_comm* comm, int sm_count, cudaStream_t st, std::vector<double>& mean_host,
                          std::vector<double>& k_host, double t_total, picard_stats_t* stats) {
  const bool tr = getenv("PICARD_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!tr) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[picard trace]   %-26s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t_prev).count());
    t_prev = t1;
  };
  mean_host.clear();
  k_host.clear();
  DevBuf<double> dmean((size_t)nf);
  dmean.zero(st);
  if (centering) {  // center: whitening.rs:24-35 (row means; the subtraction is folded into the passes as a bias)
    DevBuf<double> work((size_t)nf * 1024);
    stats->kernel_launches += aux::row_sums(d_x, nf, t_local, ldx, work.p, dmean.p, st);
    comm_allreduce_sum(comm, dmean.p, (size_t)nf, st);
    mean_host.resize((size_t)nf);
    PICARD_CUDA(cudaMemcpyAsync(mean_host.data(), dmean.p, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nf; ++i) mean_host[i] /= t_total;
    PICARD_CUDA(cudaMemcpyAsync(dmean.p, mean_host.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, st));
  }
  mark("centering");
  if (!whiten) return;
  if (nc > nf)  // whitening.rs:51-58
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: n_components (" + std::to_string(nc) + ") cannot exceed n_features (" +
                                               std::to_string(nf) + ")");
  // whiten: whitening.rs:48-116.  The thin SVD of the N x T matrix (dgesvd, U and s only) is replaced by the
  // symmetric eigenproblem of X_c X_c^T = U S^2 U^T: a SYRK-shaped pass (the moments kernel with psi(y) = y,
  // W = I, bias = mean) + one allreduce + a single-CTA Jacobi eigensolver.
  const size_t nn = (size_t)nf * nf;
  DevBuf<double> eye(nn), mom((size_t)mom_size(nf) + MOM_EXTRA), partial(pass_workspace_doubles(nf, sm_count)), V(nn), ev((size_t)nf),
      eig_scratch(nn + 8);
  mark("whiten: allocations");
  stats->kernel_launches += small::set_identity(eye.p, nf, st);
  PassLaunch L;
  L.d_x = d_x; L.ldx = ldx; L.t_local = t_local; L.n_in = nf; L.n_out = nf; L.d_w = eye.p; L.ldw = nf;
  L.d_bias = centering ? dmean.p : nullptr; L.dens = DENS_LINEAR; L.alpha = 1.0; L.mode = PASS_GRAD; L.want_h = false;
  L.d_partial = partial.p; L.d_mom = mom.p; L.d_out = nullptr; L.ld_out = 0; L.sm_count = sm_count; L.stream = st;
  stats->kernel_launches += launch_pass(L);
  mark("whiten: covariance pass");
  comm_allreduce_sum(comm, mom.p + mom_off_gr(nf), nn, st);
  stats->kernel_launches += small::jacobi_eigh(mom.p + mom_off_gr(nf), nf, V.p, ev.p, eig_scratch.p, st);
  mark("whiten: eigensolver");
  std::vector<double> evals((size_t)nf), U(nn);
  PICARD_CUDA(cudaMemcpyAsync(evals.data(), ev.p, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaMemcpyAsync(U.data(), V.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaStreamSynchronize(st));
  // singular values descending = sqrt of eigenvalues descending (dgesvd order).
  // The eigenvalues of the Gram matrix carry rounding noise ~eps * lambda_max, so singular values below ~1e-3 sigma_max come out
  // with a relative error above 1e-10 and those below ~1e-7 sigma_max cannot be told from zero, while the reference's SVD of X
  // itself resolves them down to its absolute threshold of 1e-10 (whitening.rs:61-79).  When the kept spectrum reaches that
  // range the decomposition is REFINED (two more N x T passes, rare path):
  //   K1 = S1^-1 U^T from the first eigh (tiny / negative eigenvalues floored) ; C2 = K1 C K1^T accumulated from the data in f64
  //   (= I where the first stage was right) ; C2 = V D V^T (host Jacobi) ; then C = M M^T with M = U S1 V D^1/2, and the SVD of M --
  //   one-sided Jacobi on (D^1/2 V^T) S1, a well-conditioned matrix times a graded diagonal: high RELATIVE accuracy for the small
  //   singular values -- gives U and sigma of X_c as an SVD of X_c would.
  const double lam_max = std::fmax(evals[nf - 1], 0.0);
  if (nc >= 1 && evals[nf - nc] < 1e-6 * lam_max && lam_max > 0.0) {
    const double floor_l = 4.0 * 2.220446049250313e-16 * lam_max;
    std::vector<double> s1((size_t)nf), k1(nn);
    for (int i = 0; i < nf; ++i) {
      s1[i] = std::sqrt(std::fmax(evals[i], floor_l));
      for (int j = 0; j < nf; ++j) k1[(size_t)i * nf + j] = U[(size_t)j * nf + i] / s1[i];
    }
    const int64_t ldz = round_up(t_local, 16);
    DevBuf<double> z((size_t)nf * ldz);
    stats->kernel_launches += apply_device(k1.data(), centering ? mean_host.data() : nullptr, nf, nf, d_x, ldx, z.p, ldz, t_local, sm_count, st);
    PassLaunch L2 = L;
    L2.d_x = z.p; L2.ldx = ldz; L2.d_bias = nullptr;
    stats->kernel_launches += launch_pass(L2);
    comm_allreduce_sum(comm, mom.p + mom_off_gr(nf), nn, st);
    std::vector<double> c2(nn), vd, V2;
    PICARD_CUDA(cudaMemcpyAsync(c2.data(), mom.p + mom_off_gr(nf), sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nf; ++i)
      for (int j = i + 1; j < nf; ++j) { const double a = 0.5 * (c2[(size_t)i * nf + j] + c2[(size_t)j * nf + i]); c2[(size_t)i * nf + j] = c2[(size_t)j * nf + i] = a; }
    host_jacobi_eigh(c2, nf, vd, V2);
    // B = (D^1/2 V^T) S1 : b[i][j] = sqrt(d_i) V2[j][i] s1[j]
    std::vector<double> b(nn), jrot;
    for (int i = 0; i < nf; ++i) {
      const double sd_i = std::sqrt(std::fmax(vd[i], 0.0));
      for (int j = 0; j < nf; ++j) b[(size_t)i * nf + j] = sd_i * V2[(size_t)j * nf + i] * s1[j];
    }
    host_one_sided_jacobi(b, nf, jrot);
    // sigma_j^2 = squared norm of column j of the rotated B ; left singular vectors of X_c = U jrot ; sort ascending like eigh
    std::vector<double> lam((size_t)nf), Unew(nn);
    for (int j = 0; j < nf; ++j) { double sj = 0; for (int k = 0; k < nf; ++k) sj += b[(size_t)k * nf + j] * b[(size_t)k * nf + j]; lam[j] = sj; }
    for (int i = 0; i < nf; ++i)
      for (int j = 0; j < nf; ++j) { double acc = 0; for (int k = 0; k < nf; ++k) acc += U[(size_t)i * nf + k] * jrot[(size_t)k * nf + j]; Unew[(size_t)i * nf + j] = acc; }
    std::vector<int> order((size_t)nf);
    for (int i = 0; i < nf; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int c) { return lam[a] < lam[c]; });
    for (int jn = 0; jn < nf; ++jn) {
      evals[jn] = lam[order[jn]];
      for (int i = 0; i < nf; ++i) U[(size_t)i * nf + jn] = Unew[(size_t)i * nf + order[jn]];
    }
    mark("whiten: refinement (ill-conditioned data)");
  }
  double min_sv = INFINITY;
  for (int i = 0; i < nc; ++i) {
    const double ev = evals[nf - 1 - i];
    min_sv = std::fmin(min_sv, ev > 0.0 ? std::sqrt(ev) : 0.0);
  }
  if (!(min_sv >= 1e-10)) throw Error(PICARD_SINGULAR_MATRIX, "Singular matrix encountered during computation");  // whitening.rs:72-79
  const double scale = std::sqrt(t_total);  // whitening.rs:83
  k_host.assign((size_t)nc * nf, 0.0);
  for (int i = 0; i < nc; ++i) {
    const int col = nf - 1 - i;
    const double s = std::sqrt(evals[col]);
    for (int j = 0; j < nf; ++j) k_host[(size_t)i * nf + j] = U[(size_t)j * nf + col] / s * scale;
    // sign rule, whitening.rs:93-107 (quirk Q16; Iterator::max_by keeps the LAST maximum on ties)
    int best = 0;
    for (int j = 0; j < nf; ++j)
      if (std::fabs(k_host[(size_t)i * nf + j]) >= std::fabs(k_host[(size_t)i * nf + best])) best = j;
    if (k_host[(size_t)i * nf + best] < 0.0)
      for (int j = 0; j < nf; ++j) k_host[(size_t)i * nf + j] = -k_host[(size_t)i * nf + j];
  }
}

// Device (rows x cols, leading dimension src_ld) -> pageable host (rows x cols, contiguous).  A plain cudaMemcpy2D into
// pageable memory ran at ~2 GB/s on the B200 hosts.  Here W worker threads each run their own two-deep pipeline on their
// own stream: DMA a chunk into a pinned staging buffer, then memcpy it into the destination while the next chunk is in
// flight.  No cross-thread synchronisation; PCIe and host memory bandwidth are both kept busy.
namespace {
struct StagingArena {  // process-wide pinned staging, allocated once (page-locking is slow), reused by every call
  std::mutex mu;
  unsigned char* base = nullptr;
  size_t bytes = 0;
  bool busy = false;
};
StagingArena g_arena;
constexpr size_t kStageBytes = (size_t)16 << 20;
constexpr int kMaxWorkers = 8;
}  // namespace

static bool result_is_pinned(const double* p);
static void d2h_pipelined(double* dst, const double* src_dev, int64_t src_ld, int64_t rows, int64_t cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  const size_t row_bytes = sizeof(double) * (size_t)cols;
  const size_t total = row_bytes * (size_t)rows;
  if (total < ((size_t)32 << 20) || result_is_pinned(dst)) {  // small, or a pinned destination: one plain copy (DMA at PCIe speed)
    PICARD_CUDA(cudaMemcpy2DAsync(dst, row_bytes, src_dev, sizeof(double) * src_ld, row_bytes, rows, cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
    return;
  }
  unsigned hw = std::thread::hardware_concurrency();
  const int nworkers = (int)std::max(1u, std::min((unsigned)kMaxWorkers, hw ? hw / 2 : 2u));
  // staging: the shared arena if it is free, otherwise a private allocation (concurrent calls)
  const size_t need = (size_t)nworkers * 2 * kStageBytes;
  unsigned char* stage_base = nullptr;
  bool from_arena = false;
  {
    std::lock_guard<std::mutex> lk(g_arena.mu);
    if (!g_arena.busy) {
      if (g_arena.bytes < need) {
        if (g_arena.base) cudaFreeHost(g_arena.base);
        g_arena.base = nullptr; g_arena.bytes = 0;
        if (cudaMallocHost(&g_arena.base, need) == cudaSuccess) g_arena.bytes = need; else cudaGetLastError();
      }
      if (g_arena.base) { g_arena.busy = true; from_arena = true; stage_base = g_arena.base; }
    }
  }
  PinnedBuf<unsigned char> private_stage(from_arena ? 1 : need);
  if (!from_arena) stage_base = private_stage.p;
  struct ArenaRelease { bool on; ~ArenaRelease() { if (on) { std::lock_guard<std::mutex> lk(g_arena.mu); g_arena.busy = false; } } } rel{from_arena};

  // chunks: a block of whole rows, or a column range of one row when a row is longer than a staging buffer
  const int64_t cols_per_chunk = row_bytes > kStageBytes ? (int64_t)(kStageBytes / sizeof(double)) : cols;
  const int64_t rows_per_chunk = row_bytes > kStageBytes ? 1 : std::max<int64_t>(1, (int64_t)(kStageBytes / row_bytes));
  struct Chunk { int64_t r0, nr, c0, nc; };
  std::vector<Chunk> chunks;
  for (int64_t r = 0; r < rows; r += rows_per_chunk)
    for (int64_t c = 0; c < cols; c += cols_per_chunk)
      chunks.push_back({r, std::min(rows_per_chunk, rows - r), c, std::min(cols_per_chunk, cols - c)});
  cudaEvent_t ready;
  PICARD_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  PICARD_CUDA(cudaEventRecord(ready, st));  // the source is complete once everything queued on `st` so far has run
  int device = 0;
  PICARD_CUDA(cudaGetDevice(&device));
  std::vector<int> status((size_t)nworkers, 0);
  std::vector<std::thread> th;
  for (int w = 0; w < nworkers; ++w) {
    th.emplace_back([&, w] {
      if (cudaSetDevice(device) != cudaSuccess) { status[w] = 1; return; }
      cudaStream_t s; cudaEvent_t ev[2];
      if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { status[w] = 1; return; }
      cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
      cudaStreamWaitEvent(s, ready, 0);
      unsigned char* stg[2] = {stage_base + (size_t)(2 * w) * kStageBytes, stage_base + (size_t)(2 * w + 1) * kStageBytes};
      std::vector<size_t> mine;
      for (size_t k = (size_t)w; k < chunks.size(); k += (size_t)nworkers) mine.push_back(k);
      auto drain = [&](const Chunk& ch, const unsigned char* sbuf) {
        const size_t cb = sizeof(double) * (size_t)ch.nc;
        for (int64_t r = 0; r < ch.nr; ++r)
          memcpy(dst + (size_t)(ch.r0 + r) * (size_t)cols + (size_t)ch.c0, sbuf + (size_t)r * cb, cb);
      };
      for (size_t i = 0; i <= mine.size(); ++i) {
        if (i < mine.size()) {
          const Chunk& ch = chunks[mine[i]];
          if (cudaMemcpy2DAsync(stg[i & 1], sizeof(double) * ch.nc, src_dev + (size_t)ch.r0 * src_ld + ch.c0, sizeof(double) * src_ld,
                                sizeof(double) * ch.nc, ch.nr, cudaMemcpyDeviceToHost, s) != cudaSuccess) status[w] = 1;
          cudaEventRecord(ev[i & 1], s);
        }
        if (i > 0) {
          if (cudaEventSynchronize(ev[(i - 1) & 1]) != cudaSuccess) status[w] = 1;
          drain(chunks[mine[i - 1]], stg[(i - 1) & 1]);
        }
      }
      cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); cudaStreamDestroy(s);
    });
  }
  for (auto& t : th) t.join();
  cudaEventDestroy(ready);
  for (int w = 0; w < nworkers; ++w)
    if (status[w]) throw Error(PICARD_COMPUTATION_ERROR, std::string("Computation error: CUDA failure in the device-to-host copy: ") +
                                                             cudaGetErrorString(cudaGetLastError()));
}

// Small result buffers are released with free() (picard_result_free).  A LARGE `sources` result (>= 256 MB: 10 GB at c3) comes from
// a process-wide PINNED arena instead: the device-to-host copy then is one DMA at PCIe speed straight into the result (no staging
// memcpy, no first-touch page faults), and the time does not depend on what the host's page cache is doing.  Page-locking is slow
// (seconds for 10 GB), so the arena is allocated once, kept, and reused by later fits; while a result still owns it (or the
// allocation fails) fits fall back to malloc + the staged copy.  picard_result_free() returns it, picard_release_cache() frees it.
namespace {
struct ResultArena {
  std::mutex mu;
  double* base = nullptr;
  size_t bytes = 0;
  bool busy = false;
};
ResultArena g_result_arena;
constexpr size_t kArenaMinBytes = (size_t)256 << 20;
}  // namespace

static double* result_arena_acquire(size_t count) {
  const size_t bytes = sizeof(double) * count;
  if (bytes < kArenaMinBytes || getenv("PICARD_NO_PINNED_RESULT") != nullptr) return nullptr;
  std::lock_guard<std::mutex> lk(g_result_arena.mu);
  if (g_result_arena.busy) return nullptr;
  if (g_result_arena.bytes < bytes) {
    if (g_result_arena.base) cudaFreeHost(g_result_arena.base);
    g_result_arena.base = nullptr; g_result_arena.bytes = 0;
    const size_t want = (bytes + ((size_t)64 << 20) - 1) & ~(((size_t)64 << 20) - 1);
    const double t0 = trace_now_ms();
    if (cudaHostAlloc((void**)&g_result_arena.base, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); g_result_arena.base = nullptr; return nullptr; }
    trace_slow("cudaHostAlloc (result arena)", want, t0);
    g_result_arena.bytes = want;
  }
  g_result_arena.busy = true;
  return g_result_arena.base;
}
bool result_arena_release(double* p) {  // true: p was the arena (now free for the next fit)
  if (!p) return false;
  std::lock_guard<std::mutex> lk(g_result_arena.mu);
  if (p != g_result_arena.base) return false;
  g_result_arena.busy = false;
  return true;
}
void result_arena_free() {
  std::lock_guard<std::mutex> lk(g_result_arena.mu);
  if (g_result_arena.base && !g_result_arena.busy) { cudaFreeHost(g_result_arena.base); g_result_arena.base = nullptr; g_result_arena.bytes = 0; }
}
static bool result_is_pinned(const double* p) {
  std::lock_guard<std::mutex> lk(g_result_arena.mu);
  return p != nullptr && p == g_result_arena.base;
}
static void free_result(double* p) { if (!result_arena_release(p)) free(p); }

static double* alloc_result(size_t count) {
  if (double* a = result_arena_acquire(count)) return a;
  void* p = malloc(sizeof(double) * (count ? count : 1));
  if (!p) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: out of host memory");
  return (double*)p;
}

// The `sources` result (nc x T doubles, 10 GB at c3) is a fresh allocation whose pages are first touched by the D2H
// copy: ~2.6 M page faults that cap the copy at a few GB/s.  They are taken here instead, by background threads, while
// the GPU runs the solver; join() before the copy.
struct Prefault {
  std::vector<std::thread> th;
  void start(double* p, size_t count) {
    const size_t bytes = sizeof(double) * count;
    if (bytes < ((size_t)64 << 20) || result_is_pinned(p)) return;  // the pinned arena is resident already
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)std::max(1u, std::min(16u, hw ? hw / 2 : 2u));
    unsigned char* base = reinterpret_cast<unsigned char*>(p);
    const size_t per = ((bytes / n) + 4095) & ~(size_t)4095;
    for (int t = 0; t < n; ++t) {
      const size_t a = (size_t)t * per, b = std::min(bytes, a + per);
      if (a >= b) break;
      th.emplace_back([=] { for (size_t o = a; o < b; o += 4096) base[o] = 0; });
    }
  }
  void join() { for (auto& t : th) if (t.joinable()) t.join(); th.clear(); }
  ~Prefault() { join(); }
};

// fit_host knows the shape of `sources` before the H2D copy starts: it allocates the result and starts the page faults there,
// so that they also overlap the upload and the preprocessing; fit_device adopts the buffer (same thread) if the size matches.
struct PendingResult {
  double* p = nullptr;
  size_t count = 0;
  Prefault pf;
  void drop() { pf.join(); free_result(p); p = nullptr; count = 0; }
};
static thread_local PendingResult* g_pending_result = nullptr;

static double* dup_host(const double* p, size_t n) {
  double* o = (double*)malloc(sizeof(double) * (n ? n : 1));
  if (!o) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: out of host memory");
  memcpy(o, p, sizeof(double) * n);
  return o;
}

namespace {
// PICARD_TRACE=1: wall-clock stage timings of the fit pipeline on stderr (diagnostics only)
struct Trace {
  bool on = getenv("PICARD_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what, cudaStream_t st) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[picard trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

void fit_device(const double* d_x, int64_t n_features, int64_t n_samples, int64_t ldx, const picard_config_t& cfg,
                double* d_sources, int64_t lds, picard_result_t* out) {
  memset(out, 0, sizeof *out);
  Trace trace;
  config_validate(cfg);                                                       // solver.rs:46
  if (n_features <= 0 || n_samples <= 0)                                      // solver.rs:50-54
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: Input matrix cannot be empty");
  DeviceGuard guard(cfg.device);
  cudaStream_t st;
  PICARD_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamDel { cudaStream_t s; ~StreamDel() { cudaStreamDestroy(s); } } sdel{st};
  cudaEvent_t e0, e1;
  PICARD_CUDA(cudaEventCreate(&e0)); PICARD_CUDA(cudaEventCreate(&e1));
  struct EvDel { cudaEvent_t a, b; ~EvDel() { cudaEventDestroy(a); cudaEventDestroy(b); } } edel{e0, e1};
  PICARD_CUDA(cudaEventRecord(e0, st));
  picard_stats_t stats;
  memset(&stats, 0, sizeof stats);

  const int nf = (int)n_features;
  const int64_t t_local = n_samples;
  double t_total = (double)t_local;
  if (cfg.comm && comm_size(cfg.comm) > 1) {
    DevBuf<double> tmp(1);
    PICARD_CUDA(cudaMemcpyAsync(tmp.p, &t_total, sizeof(double), cudaMemcpyHostToDevice, st));
    comm_allreduce_sum(cfg.comm, tmp.p, 1, st);
    PICARD_CUDA(cudaMemcpyAsync(&t_total, tmp.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
  }
  const int64_t mn = std::min<int64_t>(nf, (int64_t)t_total);
  int64_t ncomp = cfg.n_components >= 0 ? cfg.n_components : mn;              // solver.rs:63
  if (ncomp > mn) ncomp = mn;
  const bool ortho = cfg.ortho != 0;
  const bool extended = cfg.extended < 0 ? ortho : (cfg.extended != 0);      // solver.rs:66
  if (cfg.density_kind != PICARD_DENSITY_TANH && extended && !ortho && (!cfg.comm || comm_rank(cfg.comm) == 0))  // solver.rs:69-74
    fprintf(stderr, "Warning: Using a density other than tanh with extended=true and ortho=false may result in incorrect estimation or numerical overflow\n");

  std::vector<double> mean, K;
  center_whiten_device(d_x, nf, t_local, ldx, (int)ncomp, cfg.centering != 0, cfg.whiten != 0, cfg.comm, guard.sm_count, st, mean, K,
                       t_total, &stats);                                      // solver.rs:77-93
  trace.mark("center + whiten", st);
  const int nc = cfg.whiten ? (int)ncomp : nf;                                // solver.rs:95 (quirk Q13)
  pass_padded_size(nc);

  // w_init (solver.rs:98-121)
  std::vector<double> w_init((size_t)nc * nc);
  if (cfg.w_init) {
    if ((cfg.w_init_rows || cfg.w_init_cols) && (cfg.w_init_rows != nc || cfg.w_init_cols != nc))
      throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: w_init shape [" + std::to_string(cfg.w_init_rows) + ", " +
                                                 std::to_string(cfg.w_init_cols) + "] doesn't match expected (" + std::to_string(nc) +
                                                 ", " + std::to_string(nc) + ")");
    memcpy(w_init.data(), cfg.w_init, sizeof(double) * nc * nc);
  } else {
    uint64_t seed = cfg.seed;
    if (!cfg.has_seed) { std::random_device rd; seed = ((uint64_t)rd() << 32) ^ rd(); }
    std::vector<double> g((size_t)nc * nc);
    randn_fill(seed, g.data(), g.size());
    DevBuf<double> dg((size_t)nc * nc), dwork(small::sym_decorrelation_work(nc)), dout((size_t)nc * nc);
    DevBuf<int> dst(1);
    PICARD_CUDA(cudaMemcpyAsync(dg.p, g.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, st));
    stats.kernel_launches += small::sym_decorrelation(dg.p, nc, dwork.p, dout.p, dst.p, st);
    int status = 0;
    PICARD_CUDA(cudaMemcpyAsync(&status, dst.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaMemcpyAsync(w_init.data(), dout.p, sizeof(double) * nc * nc, cudaMemcpyDeviceToHost, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
    if (status != PICARD_OK) throw Error(PICARD_SINGULAR_MATRIX, "Singular matrix encountered during computation");
  }

  // x1 = K (x - mean) when a warm start needs the whitened data itself (solver.rs:124-137)
  trace.mark("w_init", st);
  const int64_t ld1 = round_up(t_local, 16);
  DevBuf<double> x1((size_t)nc * ld1);
  trace.mark("alloc x1", st);
  std::vector<double> eye_nc;
  if (cfg.jade_it >= 0 || cfg.fastica_it >= 0) {
    const bool is_jade = cfg.jade_it >= 0;  // JADE takes priority (solver.rs:124-137; validate() forbids both)
    if (cfg.verbose && (!cfg.comm || comm_rank(cfg.comm) == 0))
      printf("Running %lld iterations of %s...\n", (long long)(is_jade ? cfg.jade_it : cfg.fastica_it), is_jade ? "JADE" : "FastICA");
    std::vector<double> a0;
    const double* a_ptr;
    if (cfg.whiten) a_ptr = K.data();
    else { a0.assign((size_t)nc * nc, 0.0); for (int i = 0; i < nc; ++i) a0[(size_t)i * nc + i] = 1.0; a_ptr = a0.data(); }
    stats.kernel_launches += apply_device(a_ptr, mean.empty() ? nullptr : mean.data(), nc, nf, d_x, ldx, x1.p, ld1, t_local,
                                          guard.sm_count, st);
    if (is_jade) {
      jade_device(x1.p, nc, t_local, ld1, t_total, cfg.jade_it, 1e-6, cfg.verbose != 0, cfg.comm, guard.sm_count, st, w_init.data(),
                  nullptr, &stats);                                           // replaces w_init (quirk Q14)
    } else {  // ica_par(&x1, &density, fastica_it, &w_init) -- starts from w_init (solver.rs:136)
      picard_config_t fc = cfg;
      fc.ortho = 1;  // gradient moments only (no H)
      if (pass_padded_size(nc) <= 128) fc.flags |= PICARD_FLAG_NO_Y_STORE;  // the from-X gradient pass is used: no Y buffer needed
      CoreSolver fica(x1.p, nc, t_local, ld1, fc, false, guard.sm_count, st);
      fica.fastica(cfg.fastica_it, w_init.data());
      stats.kernel_launches += fica.stats().kernel_launches;
      if (cfg.verbose && (!cfg.comm || comm_rank(cfg.comm) == 0)) printf("FastICA pre-iterations complete.\n");
    }
  }

  // x1 = w_init * K * (x - mean)  (solver.rs:140 with whitening.rs:110 folded in)
  std::vector<double> a_total((size_t)nc * nf);
  if (cfg.whiten) host_matmul(w_init.data(), K.data(), a_total.data(), nc, nc, nf);
  else a_total = w_init;
  stats.kernel_launches += apply_device(a_total.data(), mean.empty() ? nullptr : mean.data(), nc, nf, d_x, ldx, x1.p, ld1, t_local,
                                        guard.sm_count, st);
  trace.mark("x1 = w_init K (x - mean)", st);
  PICARD_CUDA(cudaEventRecord(e1, st));
  PICARD_CUDA(cudaStreamSynchronize(st));
  float pre_ms = 0.f;
  PICARD_CUDA(cudaEventElapsedTime(&pre_ms, e0, e1));
  stats.preprocess_ms = pre_ms;

  const bool keep_dev = (cfg.flags & PICARD_FLAG_KEEP_SOURCES_ON_DEVICE) != 0;
  struct HostBuf { double* p = nullptr; ~HostBuf() { free_result(p); } double* release() { double* q = p; p = nullptr; return q; } } src_guard;
  Prefault prefault;
  if (!keep_dev) {
    PendingResult* pr = g_pending_result;
    if (pr && pr->p && pr->count == (size_t)nc * t_local) {  // allocated (and being faulted in) since before the upload
      src_guard.p = pr->p; pr->p = nullptr; pr->count = 0;
      prefault.th = std::move(pr->pf.th);
    } else {
      src_guard.p = alloc_result((size_t)nc * t_local);
      prefault.start(src_guard.p, (size_t)nc * t_local);
    }
  }
  if (cfg.verbose && (!cfg.comm || comm_rank(cfg.comm) == 0)) printf("Running Picard...\n");
  CoreSolver core(x1.p, nc, t_local, ld1, cfg, extended && cfg.whiten, guard.sm_count, st);  // solver.rs:143-166
  trace.mark("core solver setup", st);
  core.run(cfg.max_iter);
  trace.mark("core loop", st);
  std::vector<double> wc((size_t)nc * nc), signs((size_t)nc);
  core.state(wc.data(), signs.data(), nullptr, nullptr, nullptr, nullptr);
  std::vector<double> wfull((size_t)nc * nc);
  host_matmul(wc.data(), w_init.data(), wfull.data(), nc, nc, nc);            // solver.rs:169
  if (!core.converged() && cfg.verbose && (!cfg.comm || comm_rank(cfg.comm) == 0))
    fprintf(stderr, "Warning: PICARD did not converge. Final gradient norm: %.4e, tolerance: %.4e\n", core.gradient_norm(), cfg.tol);

  // sources = Y = W_core x1 (the reference returns the Y the norm was measured at: quirk Q10)
  const picard_stats_t& cs = core.stats();
  stats.core_ms = cs.core_ms; stats.fused_passes = cs.fused_passes; stats.grad_passes = cs.grad_passes; stats.loss_passes = cs.loss_passes;
  stats.ls_tries = cs.ls_tries; stats.fallbacks = cs.fallbacks; stats.sign_changes = cs.sign_changes;
  stats.kernel_launches += cs.kernel_launches;
  stats.pass_ms_fused = cs.pass_ms_fused; stats.pass_ms_grad = cs.pass_ms_grad; stats.pass_ms_loss = cs.pass_ms_loss;
  stats.grady_passes = cs.grady_passes; stats.pass_ms_grady = cs.pass_ms_grady;
  stats.i8_loss_passes = cs.i8_loss_passes; stats.i8_grad_passes = cs.i8_grad_passes; stats.i8_fallbacks = cs.i8_fallbacks;
  stats.i8_range = cs.i8_range;
  double* host_sources = nullptr;
  prefault.join();
  trace.mark("  sources: result buffer ready", st);
  if (d_sources) {
    stats.kernel_launches += apply_device(wc.data(), nullptr, nc, nc, x1.p, ld1, d_sources, lds, t_local, guard.sm_count, st);
    if (!keep_dev) {
      d2h_pipelined(src_guard.p, d_sources, lds, nc, t_local, st);
      host_sources = src_guard.release(); out->sources = host_sources;
      stats.d2h_bytes += (int64_t)sizeof(double) * nc * t_local;
    }
  } else if (!keep_dev) {
    core.release_pass_buffers();  // Y store + digit image of x1: not needed any more, and `ysrc` is as large as either
    trace.mark("  sources: pass buffers released", st);
    DevBuf<double> ysrc((size_t)nc * ld1);
    stats.kernel_launches += apply_device(wc.data(), nullptr, nc, nc, x1.p, ld1, ysrc.p, ld1, t_local, guard.sm_count, st);
    trace.mark("  sources: Y = W x1", st);
    cudaEvent_t d0, d1;
    PICARD_CUDA(cudaEventCreate(&d0)); PICARD_CUDA(cudaEventCreate(&d1));
    PICARD_CUDA(cudaEventRecord(d0, st));
    d2h_pipelined(src_guard.p, ysrc.p, ld1, nc, t_local, st);
    host_sources = src_guard.release(); out->sources = host_sources;
    PICARD_CUDA(cudaEventRecord(d1, st));
    PICARD_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; cudaEventElapsedTime(&ms, d0, d1); stats.d2h_ms += ms;
    cudaEventDestroy(d0); cudaEventDestroy(d1);
    stats.d2h_bytes += (int64_t)sizeof(double) * nc * t_local;
    trace.mark("  sources: D2H", st);
  }

  trace.mark("sources: buffers freed", st);
  out->n_components = nc; out->n_features = nf; out->n_samples = t_local;
  out->whitening = cfg.whiten ? dup_host(K.data(), K.size()) : nullptr;
  out->unmixing = dup_host(wfull.data(), wfull.size());
  out->sources = host_sources;
  out->mean = mean.empty() ? nullptr : dup_host(mean.data(), mean.size());
  out->n_iterations = core.n_iterations(); out->converged = core.converged() ? 1 : 0; out->gradient_norm = core.gradient_norm();
  out->signs = core.extended() ? dup_host(signs.data(), signs.size()) : nullptr;
  out->stats = stats;
}

void fit_host(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, const picard_config_t& cfg,
              picard_result_t* out) {
  memset(out, 0, sizeof *out);
  config_validate(cfg);
  if (n_features <= 0 || n_samples <= 0 || x == nullptr)
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: Input matrix cannot be empty");
  DeviceGuard guard(cfg.device);
  // shape of the result when it does not depend on the other ranks: nc = n_components (whitening) or n_features
  PendingResult pending;
  struct PendingGuard { PendingResult* r; ~PendingGuard() { g_pending_result = nullptr; r->drop(); } } pending_guard{&pending};
  if (n_features <= n_samples) {
    const int64_t ncomp = cfg.n_components >= 0 ? std::min<int64_t>(cfg.n_components, n_features) : n_features;
    const int64_t nc = cfg.whiten ? ncomp : n_features;
    if (nc > 0) {
      pending.count = (size_t)nc * (size_t)n_samples;
      pending.p = alloc_result(pending.count);
      pending.pf.start(pending.p, pending.count);
      g_pending_result = &pending;
    }
  }
  const int64_t ldx = round_up(n_samples, 16);
  DevBuf<double> dx((size_t)n_features * ldx);
  cudaEvent_t e0, e1;
  PICARD_CUDA(cudaEventCreate(&e0)); PICARD_CUDA(cudaEventCreate(&e1));
  PICARD_CUDA(cudaEventRecord(e0, 0));
  PICARD_CUDA(cudaMemcpy2DAsync(dx.p, sizeof(double) * ldx, x, sizeof(double) * row_stride, sizeof(double) * n_samples, n_features,
                                cudaMemcpyHostToDevice, 0));
  PICARD_CUDA(cudaEventRecord(e1, 0));
  PICARD_CUDA(cudaStreamSynchronize(0));
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  picard_config_t c2 = cfg;
  c2.device = guard.device;
  c2.flags &= ~PICARD_FLAG_KEEP_SOURCES_ON_DEVICE;
  fit_device(dx.p, n_features, n_samples, ldx, c2, nullptr, 0, out);
  out->stats.h2d_ms += ms;
  out->stats.h2d_bytes += (int64_t)sizeof(double) * n_features * n_samples;
}

void transform_host(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, const picard_result_t& res,
                    double* out, int device) {
  if (n_features <= 0 || n_samples <= 0) throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: Input matrix cannot be empty");
  if (n_features != res.n_features)
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: transform input has " + std::to_string(n_features) +
                                               " features, the model was fitted on " + std::to_string(res.n_features));
  DeviceGuard guard(device);
  const int nc = (int)res.n_components, nf = (int)res.n_features;
  std::vector<double> wfull((size_t)nc * nf);
  if (res.whitening) host_matmul(res.unmixing, res.whitening, wfull.data(), nc, nc, nf);  // result.rs:39-44
  else memcpy(wfull.data(), res.unmixing, sizeof(double) * nc * nc);
  const int64_t ldx = round_up(n_samples, 16);
  DevBuf<double> dx((size_t)nf * ldx), dy((size_t)nc * ldx);
  PICARD_CUDA(cudaMemcpy2DAsync(dx.p, sizeof(double) * ldx, x, sizeof(double) * row_stride, sizeof(double) * n_samples, nf,
                                cudaMemcpyHostToDevice, 0));
  apply_device(wfull.data(), res.mean, nc, nf, dx.p, ldx, dy.p, ldx, n_samples, guard.sm_count, 0);
  d2h_pipelined(out, dy.p, ldx, nc, n_samples, 0);
}

}  // namespace picard
