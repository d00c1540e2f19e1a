// Host-side engine of libpicard_b200.so: device buffers, the resumable core solver (core.rs:162-401 control
// flow, device-resident state), the fit pipeline (solver.rs:45-189) and the sample-axis communicator.
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "pass.cuh"
#include "small.cuh"

struct picard_comm;

namespace picard {

// ---- communicator (comm.cu)
void comm_unique_id(char id[PICARD_UNIQUE_ID_BYTES]);
picard_comm* comm_create(const char id[PICARD_UNIQUE_ID_BYTES], int rank, int nranks, int device);
void comm_destroy(picard_comm* c);
int comm_rank(const picard_comm* c);
int comm_size(const picard_comm* c);
void comm_allreduce_sum(picard_comm* c, double* d_buf, size_t count, cudaStream_t st);
void comm_allreduce_sum2(picard_comm* c, double* a, size_t na, double* b, size_t nb, cudaStream_t st);
struct P2PCall;
// peer-memory exchange carried by a pass kernel's own tail: false = not available (single rank, no peer access, payload too large)
bool comm_p2p_next(picard_comm* c, size_t count, P2PCall* out);

// PICARD_TRACE diagnostics: report driver allocator calls that take more than 5 ms
double trace_now_ms();
void trace_slow(const char* what, size_t bytes, double t0_ms);
// Device memory for the library's temporaries.  cudaMalloc / cudaFree on the B200 hosts were measured at up to 0.4 s per
// call after an idle period (50x the whitening kernels they bracket), and returning one N x T buffer to the driver at 5 - 700 ms,
// so freed blocks of every size are kept in a per-device cache and reused by the next fit (limit: PICARD_CACHE_MAX_GB, default 60 %
// of the device's memory; handed back by picard_release_cache(), and automatically when an allocation fails).
// dev_free synchronises the device first, like cudaFree does, so a block is never reused while work on it is in flight.
void* dev_alloc(size_t bytes, size_t* capacity, int* device);
void dev_free(void* p, size_t capacity, int device);  // device: the one the block was allocated on (the cache key)
void dev_cache_release();  // return every cached block to the driver
// pinned arena of large `sources` results (fit.cu)
bool result_arena_release(double* p);  // true if p was the arena: it is free for the next fit
void result_arena_free();

// ---- RAII device / pinned buffers
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap), dev(o.dev) { o.p = nullptr; o.n = 0; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; dev = o.dev; o.p = nullptr; o.n = 0; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  size_t cap = 0;  // bytes of the underlying block (>= sizeof(T) * n: blocks come from the cache below)
  int dev = 0;     // the device that owns the block
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T*>(dev_alloc(sizeof(T) * count, &cap, &dev));
  }
  void release() {
    if (p) dev_free(p, cap, dev);
    p = nullptr; n = 0; cap = 0;
  }
  void zero(cudaStream_t st) { if (p) PICARD_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * n, st)); }
};
template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  PinnedBuf() {}
  explicit PinnedBuf(size_t count) { PICARD_CUDA(cudaMallocHost(&p, sizeof(T) * (count ? count : 1))); }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

struct DeviceGuard {  // select cfg->device for the duration of a call, restore afterwards
  int prev = -1;
  explicit DeviceGuard(int device);
  ~DeviceGuard();
  int device = 0;
  int sm_count = 0;
};

// ---- auxiliary kernels (aux.cu)
namespace aux {
// row sums of (n x t_local) X -> d_out (n), deterministic two-stage reduction. work >= n * 1024 doubles.
int row_sums(const double* d_x, int n, int64_t t_local, int64_t ldx, double* d_work, double* d_out, cudaStream_t st);
int synth_sources(double* d_out, int n, int64_t t_local, int64_t ld, int64_t t_offset, int n_laplace, uint64_t seed, cudaStream_t st);
// measured FP64 tensor (DMMA) peak of the current device at its current clocks, TFLOP/s; runs for about budget_ms
double fp64_peak_probe(int sm_count, double budget_ms, cudaStream_t st);
}  // namespace aux

// ---- the core loop (core.rs:162-401)
class CoreSolver {
 public:
  CoreSolver(const double* d_x, int n, int64_t t_local, int64_t ldx, const picard_config_t& cfg, bool covariance_identity,
             int sm_count, cudaStream_t stream);
  ~CoreSolver();
  void reset();
  // Runs up to max_new further outer iterations. Returns the number performed in this call.
  int64_t run(int64_t max_new);
  void state(double* w, double* signs, int64_t* n_iterations, int32_t* converged, double* gradient_norm, double* loss);
  // FastICA parallel iterations (ica_par, solver.rs:218-249) on this solver's data: w (n x n, host) in / out
  void fastica(int64_t iters, double* w_host);
  const double* d_w() const { return W_; }
  // Frees the N x T work buffers of the passes (Y store, digit image of x1) once the loop is finished: the caller is about to
  // allocate the `sources` buffer and should not need 4 x the input in device memory at that moment.
  void release_pass_buffers() { ybuf_.release(); ybuf_valid_ = false; xs8_.release(); if (i8_state_ == 1) i8_state_ = 0; }
  const picard_stats_t& stats() const { return stats_; }
  bool converged() const { return converged_; }
  int64_t n_iterations() const { return n_iterations_; }
  double gradient_norm() const { return gradient_norm_; }
  bool extended() const { return dims_.extended != 0; }
  int n() const { return dims_.n; }

  // one evaluation of the pass at an arbitrary W into a moment buffer (test hook + internal use)
  // Returns true when the loss of the point was computed and published inside the pass kernel (finish_which >= 0 asked for it
  // and the engine supports it): the caller then skips loss_from_moments and goes straight to fetch_scalars().
  bool eval_pass(const double* d_w, int mode, bool want_h, int dens, double alpha, double* d_mom, bool store_y = false,
                 int finish_which = -1, const double* finish_signs = nullptr);
  // test hooks (picard_eval_moments / picard_eval_point): host in, host out
  void hook_moments(const double* w_host, int mode, bool want_h, double* gr, double* sd, double* hr, double* sq, double* lrow);
  void hook_point(const double* w_host, const double* c_host, const double* old_signs_host, const double* loss_signs_host,
                  double* g, double* h, double* hoff, double* signs, int32_t* sign_change, double* gradient_norm, double* loss);

 private:
  bool pass(const double* d_w, int mode, double* d_mom, int finish_which = -1, const double* finish_signs = nullptr);
  void fetch_scalars();
  void resolve_pass_time();
  void try_point(double alpha, bool speculate, int try_index, int tries_planned);

  CoreDims dims_;
  picard_config_t cfg_;
  const double* d_x_;
  int64_t t_local_, ldx_;
  int sm_count_;
  cudaStream_t st_;
  picard_comm* comm_;
  bool cov_identity_;
  int dens_;
  double alpha_;
  bool need_h_;

  DevBuf<double> store_;   // all N x N state in one allocation
  DevBuf<double> partial_;
  DevBuf<double> wt_all_;  // W' of every candidate step of the current line search (one Taylor run, small::matrix_exp_candidates)
  int cand_ready_ = 0;
  double* w_try_ = nullptr;  // W' of the last evaluated try
  DevBuf<double> ybuf_;    // Y' of the last loss-only try (n x ldx_), empty when PICARD_FLAG_NO_Y_STORE or out of memory
  bool ybuf_valid_ = false;
  int64_t ldy_ = 0;        // leading dimension of ybuf_
  // INT8 tensor-core passes (i8_loss.cu, i8_grad.cu): decided once per solver at the first LOSS pass (i8_prepare)
  DevBuf<uint8_t> xs8_;           // digit image of x1
  DevBuf<unsigned int> i8_counter_;  // [0] LOSS pass, [1] gradient pass, [2] gradient pass (exchange): CTAs counted by the kernels' tails
  unsigned int i8_counter_total_[3] = {0, 0, 0};
  DevBuf<double> xstats_;         // statistics of x1 gathered while slicing (I8_XSTATS)
  DevBuf<int> rowexp_;            // exponents of the rows of the stored Y (gradient pass)
  int i8_state_ = 0;              // 0 = undecided, 1 = in use, -1 = not used
  bool i8_prepare();
  DevBuf<CoreScalars> sc_dev_;
  PinnedBuf<CoreScalars> sc_host_;   // pinned mirror, also written by the device (publish_scalars)
  unsigned long long seq_ = 0;        // sequence number of the last publishing kernel launched
  unsigned long long next_seq() { return ++seq_; }
  // views into store_
  double *W_, *Wt_, *M_, *D_, *C_, *G_, *Gtmp_, *Gold_, *H_, *hoff_, *signs_, *old_signs_, *Sprev_, *q_;
  double *mem_s_, *mem_y_, *mem_r_;
  double *mom_cur_, *mom_trial_;
  double *lu_work_;
  small::ExpmWork ew_;

  // host-side loop state
  int64_t iter_ = 0, n_iterations_ = 0;
  bool converged_ = false, started_ = false, have_cur_ = false, speculate_next_ = true;
  double gradient_norm_ = 1.0, current_loss_ = 0.0;
  picard_stats_t stats_;
  cudaEvent_t ev_a_, ev_b_, ev_run0_, ev_run1_;
  int last_pass_mode_ = -1;
};

// ---- fit pipeline (solver.rs:45-189), fit.cu
void config_default(picard_config_t* cfg);
void config_validate(const picard_config_t& cfg);  // throws Error(PICARD_INVALID_CONFIG)
void fit_device(const double* d_x, int64_t n_features, int64_t n_samples, int64_t ldx, const picard_config_t& cfg,
                double* d_sources, int64_t lds, picard_result_t* out);
void fit_host(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, const picard_config_t& cfg,
              picard_result_t* out);
void transform_host(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, const picard_result_t& res,
                    double* out, int device);
// d_out = A (n_out x n_in, host) * (d_in - mean): the APPLY pass
int apply_device(const double* a_host, const double* mean_host, int n_out, int n_in, const double* d_in, int64_t ld_in,
                 double* d_out, int64_t ld_out, int64_t t_local, int sm_count, cudaStream_t st);
// centering + whitening on the device (whitening.rs:24-116). mean_host (nf) and k_host (nc x nf) receive the results.
void center_whiten_device(const double* d_x, int nf, int64_t t_local, int64_t ldx, int nc, bool centering, bool whiten,
                          picard_comm* comm, int sm_count, cudaStream_t st, std::vector<double>& mean_host,
                          std::vector<double>& k_host, double t_total, picard_stats_t* stats);
// the build's own documented N(0,1) generator for the random w_init path (splitmix64 + Box-Muller)
void randn_fill(uint64_t seed, double* out, size_t count);

}  // namespace picard
