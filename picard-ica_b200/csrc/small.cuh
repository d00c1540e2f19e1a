// N x N device kernels of the core loop (K3-K5, K7 of SURVEY.md §2.4): everything between two passes
// over the sample matrix runs on the device from device-resident state; the host only reads back a few
// scalars per iteration to take the (inherently sequential) convergence / line-search decisions.
#pragma once
#include "common.cuh"

namespace picard {

// Device-resident scalars of the core loop. One instance lives in device memory, a pinned host mirror is
// refreshed by cudaMemcpyAsync when the host has to decide something.
struct CoreScalars {
  double gradient_norm;   // max |G| of the projected gradient (core.rs:289)
  double current_loss;    // loss at the current iterate with the current signs
  double new_loss;        // loss of the last line-search try (1e15 if singular: core.rs:90-96)
  double norm_d;          // max |D| of the last direction (for matrix_exp's scaling, math.rs:42-49)
  double last_r;          // last 1/<s,y> (diagnostics)
  int32_t sign_change;    // core.rs:234-236
  int32_t accept;         // new_loss < current_loss
  int32_t mem_len, mem_head;  // L-BFGS ring (oldest at mem_head)
  int32_t have_g_old, have_prev_step;
  int32_t loss_singular;  // the initial / recomputed loss hit a singular W (core.rs:188-190, 321-325)
  int32_t pad;
  unsigned long long seq; // host mirror only: written LAST by the publishing kernel (see publish_scalars)
};

#ifdef __CUDACC__
// The host decides after every line-search try and every iteration front.  Instead of a D2H copy + stream synchronisation
// (~15 us of latency each) the kernel that finalises the scalars writes them into the host's pinned mirror (mapped into the
// device address space) and then, after a system-scope fence, a sequence number the host is spinning on.
__device__ __forceinline__ void publish_scalars(const CoreScalars* sc, CoreScalars* map, unsigned long long seq) {
  if (map == nullptr) return;
  const volatile CoreScalars* s = sc;
  map->gradient_norm = s->gradient_norm; map->current_loss = s->current_loss; map->new_loss = s->new_loss; map->norm_d = s->norm_d;
  map->last_r = s->last_r; map->sign_change = s->sign_change; map->accept = s->accept; map->mem_len = s->mem_len;
  map->mem_head = s->mem_head; map->have_g_old = s->have_g_old; map->have_prev_step = s->have_prev_step;
  map->loss_singular = s->loss_singular;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(&map->seq) = seq;
}
#endif

// Per-point extras appended to a moment buffer: [logdet, detsign] at offset mom_size(n).
constexpr int MOM_EXTRA = 2;

struct CoreDims {
  int n;          // components
  int m;          // L-BFGS memory size
  double t_total; // global sample count T (as f64, core.rs:176)
  int ortho, extended;
  double lambda_min;
};

namespace small {

// ---- generic helpers (each returns the number of kernels launched) ------------------------------------
// C = alpha * A * op(B) (+ I if add_identity) ; all n x n, ld = n.  trans_b: use B^T.
int matmul(const double* A, const double* B, double* C, int n, bool trans_b, double alpha, bool add_identity, cudaStream_t st);
int set_identity(double* A, int n, cudaStream_t st);
int copy_scaled(const double* A, double* B, int64_t count, double alpha, cudaStream_t st);  // B = alpha * A
int eye_plus_scaled(const double* D, double alpha, double* M, int n, cudaStream_t st);      // M = I + alpha * D (core.rs:121)

// ---- iteration front: core.rs:223-293 from reduced raw moments, + L-BFGS update (core.rs:296-331) +
// direction (lbfgs.rs:84-133), one single-CTA kernel.
struct FrontArgs {
  CoreDims d;
  const double* mom;       // reduced raw moments of the current point (+ extras)
  double* C;               // covariance-like matrix of the extended sign rule
  double* G; double* Gtmp; double* G_old; double* H; double* hoff;
  double* signs; double* old_signs;
  double* S_prev;          // last step (alpha * direction)
  double* mem_s; double* mem_y; double* mem_r;  // ring buffers [m][n*n], [m]
  double* q; double* D;    // work / output direction
  CoreScalars* sc;
  CoreScalars* sc_map = nullptr;   // host mirror to publish to (nullptr: none)
  unsigned long long seq = 0;
  int first_iter;
  int do_lbfgs;            // 1: full; 0: front only (test hook); 2: direction only from G/H/hoff/memory in place (test hook)
};
int iteration_front(const FrontArgs& a, cudaStream_t st);

// Loss of a point from its moments (core.rs:39-85 after the sums): writes sc->new_loss and sc->accept
// (which = 0) or sc->current_loss (which = 1; singular -> 1e15 and sc->loss_singular).
int loss_from_moments(const CoreDims& d, const double* mom, const double* signs, CoreScalars* sc, int which, cudaStream_t st,
                      CoreScalars* sc_map = nullptr, unsigned long long seq = 0);

// After an accepted (or forced) try: S_prev = alpha * D, current_loss = new_loss, have_prev_step = 1;
// extended with covariance = I: C = W W^T (core.rs:375-379).
int accept_step(const CoreDims& d, const double* D, double alpha, double* S_prev, const double* W, double* C, int update_c,
                CoreScalars* sc, cudaStream_t st);
int negate_into(const double* G, double* D, int64_t count, CoreScalars* sc, cudaStream_t st);  // fallback direction -G

// ---- matrix_exp (math.rs:38-74): out = expm(alpha * D). norm_d = max|D| known to the host.
struct ExpmWork { double* As; double* term0; double* term1; double* res0; double* res1; double* slots; };  // n^2 each; slots: EXPM_SLOTS
constexpr int EXPM_SLOTS = 8 + 31 * 32;
constexpr int EXPM_NC = 10;  // candidate steps per Taylor run
// W'_t = expm(alpha0 / 2^t D) W for t = 0 .. n_cand-1 (<= 10) from one Taylor run (bit-identical to n_cand calls of matrix_exp).
// Returns the number of candidates produced into Wt_all ([t][n][n]), or -1 when the case needs the per-try path
// (max |alpha0 D| > 1, i.e. squarings, or a degenerate norm).
int matrix_exp_candidates(const double* D, double alpha0, double norm_d, int n, int n_cand, const ExpmWork& w, const double* W, double* Wt_all,
                          cudaStream_t st);
// One cooperative kernel.  out (may be NULL) = expm(alpha D); if W and Wt are given, Wt = expm(alpha D) W (core.rs:125).
int matrix_exp(const double* D, double alpha, double norm_d, int n, const ExpmWork& w, double* out, cudaStream_t st,
               const double* W = nullptr, double* Wt = nullptr);

// ---- signed log-determinant by LU with partial pivoting (math.rs:84-88): out2 = [logabs, sign]
int sln_det(const double* A, int n, double* work, double* out2, cudaStream_t st);

// FastICA fixed-point matrix (solver.rs:229-238) from the pass moments at Y = W X with W orthogonal:
//   C = E[g(WX) X^T] - diag(E[g'(WX)]) W = (Gr / T - diag(Sd / T)) W     (X^T = Y^T W^-T = Y^T W)
// tmp, C: n x n.
int fastica_matrix(const double* mom, int n, double t_total, const double* W, double* tmp, double* C, cudaStream_t st);

// ---- symmetric eigendecomposition, cyclic Jacobi, single CTA (K7): A (n x n, destroyed), eigenvalues
// ascending in evals, eigenvectors in the COLUMNS of V.
// scratch: n^2 + 8 doubles.
int jacobi_eigh(double* A, int n, double* V, double* evals, double* scratch, cudaStream_t st);
// sym_decorrelation (math.rs:12-33): out = (W W^T)^{-1/2} W ; *status_dev: 0 ok, 2 singular (min eig < 1e-10)
int sym_decorrelation(const double* W, int n, double* work /* >= sym_decorrelation_work(n) doubles */, double* out, int* status_dev,
                      cudaStream_t st);
inline size_t sym_decorrelation_work(int n) { return 5 * (size_t)n * n + 2 * (size_t)n + 16; }

}  // namespace small
}  // namespace picard
