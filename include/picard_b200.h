/* =====================================================================================================
 * picard_b200.h -- C ABI of libpicard_b200.so, the B200-native (sm_100a) Picard / Picard-O ICA solver.
 *
 * Drop-in boundary for the fit path of the Rust crate lmmx/picard-ica v0.1.6.  Every entry point names the
 * reference interface it replaces (file:line under the reference's src/).  The reference-side binding a
 * maintainer would add (a Rust `extern "C"` block + safe wrapper) is shown in INTEGRATION.md; the Python
 * (ctypes) mirror lives in picard-ica_b200/ and the C++ host mirror in picard-ica_b200/host/picard.hpp.
 *
 * Conventions: all matrices are row-major f64.  Data is (n_features x n_samples): element (i, s) of `x`
 * is x[i * row_stride + s]  (ndarray `Array2<f64>` in standard layout, solver.rs:29,48).  Plain pointers
 * and sizes only; no CUDA or torch types.  `stream` arguments are a `cudaStream_t` passed as void*.
 * There is NO CPU fallback: every compute entry point fails with PICARD_COMPUTATION_ERROR if no CUDA
 * device is usable.
 * ===================================================================================================== */
#ifndef PICARD_B200_H
#define PICARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICARD_B200_ABI_VERSION 2

/* PicardError (error.rs:9-42).  NotConverged is never constructed by the reference (non-convergence
 * returns Ok{converged:false}), so it has no status code here either. */
typedef enum {
  PICARD_OK = 0,
  PICARD_INVALID_DIMENSIONS = 1, /* error.rs:22 ; solver.rs:50-54,100-108 ; whitening.rs:51-58 */
  PICARD_SINGULAR_MATRIX = 2,    /* error.rs:28 ; math.rs:22-24 ; whitening.rs:77-79 ; core.rs:188-190 */
  PICARD_COMPUTATION_ERROR = 3,  /* error.rs:31 ; LAPACK failures in the reference; CUDA/NCCL failures here */
  PICARD_INVALID_CONFIG = 4      /* error.rs:37 ; config.rs:104-142 */
} picard_status_t;

/* DensityType (density.rs:137-144): a closed enum of three kinds with one alpha parameter. */
typedef enum { PICARD_DENSITY_TANH = 0, PICARD_DENSITY_EXP = 1, PICARD_DENSITY_CUBE = 2 } picard_density_t;

/* Sample-axis communicator for multi-GPU fits (one process per GPU).  Opaque. */
typedef struct picard_comm picard_comm_t;

/* PicardConfig (config.rs:11-62), field for field; Option<T> is encoded as -1 / NULL = None.
 * Defaults (config.rs:64-85) are written by picard_config_default(). */
typedef struct {
  int32_t density_kind;   /* picard_density_t ; config.rs:13 */
  double alpha;           /* Tanh.alpha / Exp.alpha (density.rs:31-34,72-75); ignored for Cube */
  int64_t n_components;   /* config.rs:16 ; -1 = None */
  int32_t ortho;          /* config.rs:19 */
  int32_t extended;       /* config.rs:23 ; -1 = None (defaults to ortho, config.rs:99-101) */
  int32_t whiten;         /* config.rs:26 */
  int32_t centering;      /* config.rs:29 */
  int64_t max_iter;       /* config.rs:32 */
  double tol;             /* config.rs:35 */
  int64_t m;              /* config.rs:38 */
  int64_t ls_tries;       /* config.rs:41 */
  double lambda_min;      /* config.rs:44 */
  const double* w_init;   /* config.rs:47 ; (nc x nc) row-major, NULL = None */
  int64_t w_init_rows;    /* shape of w_init, checked like solver.rs:100-108 (0 = trust nc x nc) */
  int64_t w_init_cols;
  int64_t fastica_it;     /* config.rs:50 ; -1 = None */
  int64_t jade_it;        /* config.rs:55 ; -1 = None */
  int32_t has_seed;       /* config.rs:58 random_state: Option<u64> */
  uint64_t seed;
  int32_t verbose;        /* config.rs:61 */
  /* ---- execution placement: not in the reference ---- */
  int32_t device;         /* CUDA device ordinal (-1 = current device) */
  picard_comm_t* comm;    /* NULL = single GPU; otherwise `x` is this rank's column shard */
  uint32_t flags;         /* PICARD_FLAG_* */
} picard_config_t;

#define PICARD_FLAG_NO_SPECULATION 1u /* never run the speculative fused first try (it is only used when there is no Y store) */
#define PICARD_FLAG_KEEP_SOURCES_ON_DEVICE 2u /* picard_fit_device: do not copy `sources` to the host */
#define PICARD_FLAG_FORCE_SPECULATION 8u /* speculative fused first try even though the Y store is available (ablation) */
#define PICARD_FLAG_NO_Y_STORE 4u /* loss-only tries do not keep Y' (saves one N x T buffer; the gradient recomputes W X) */
/* The passes of a whitened problem with 64 < N <= 128 run on the INT8 tensor cores (error-free splitting, results within
 * ~1e-13 of the FP64 path) when a range check of the data passes; these two flags override the choice. */
#define PICARD_FLAG_NO_INT8 16u    /* FP64 (DMMA) kernels only */
#define PICARD_FLAG_FORCE_INT8 32u /* INT8 passes even if the data is not flagged as whitened / fails the range check */

/* Measurement record filled by every fit / core run (not in the reference). */
typedef struct {
  double core_ms;          /* device time of the core loop (CUDA events) */
  double preprocess_ms;    /* centering + whitening + warm start + x1 = w_init K (x - mean) */
  double h2d_ms, d2h_ms;
  int64_t h2d_bytes, d2h_bytes;
  int64_t fused_passes, grad_passes, loss_passes; /* N x T passes by kind */
  int64_t ls_tries, fallbacks, sign_changes;
  int64_t kernel_launches; /* launches of this library's own kernels */
  double pass_ms_fused, pass_ms_grad, pass_ms_loss; /* summed device time of the pass kernels by kind */
  int64_t grady_passes;    /* gradient passes served from the stored Y of an accepted loss-only try (2 N^2 T flop) */
  double pass_ms_grady;
  int64_t i8_loss_passes;  /* of loss_passes: run on the INT8 tensor cores (tcgen05.mma kind::i8) */
  int64_t i8_grad_passes;  /* of grady_passes: run on the INT8 tensor cores */
  int64_t i8_fallbacks;    /* 1 if the INT8 path was wanted but refused: range check of the data failed, or no memory for the sliced image */
  double i8_range;         /* the range check's figure: mean power-of-two sample bound / smallest row RMS of x1 (0 = not evaluated) */
} picard_stats_t;

/* PicardResult (result.rs:7-33).  Buffers are malloc'd by the library and released by
 * picard_result_free(); NULL encodes None. */
typedef struct {
  int64_t n_components, n_features, n_samples; /* n_samples = samples in THIS rank's shard */
  double* whitening;      /* result.rs:10 ; (nc x nf) or NULL */
  double* unmixing;       /* result.rs:13 ; (nc x nc) = W_core * w_init (solver.rs:169) */
  double* sources;        /* result.rs:16 ; (nc x n_samples) or NULL with PICARD_FLAG_KEEP_SOURCES_ON_DEVICE */
  double* mean;           /* result.rs:20 ; (nf) or NULL */
  int64_t n_iterations;   /* result.rs:23 */
  int32_t converged;      /* result.rs:26 */
  double gradient_norm;   /* result.rs:29 */
  double* signs;          /* result.rs:32 ; (nc), +-1.0, or NULL when not extended */
  picard_stats_t stats;
} picard_result_t;

/* ---- library ------------------------------------------------------------------------------------ */
int picard_abi_version(void);
/* Number of usable CUDA devices (0 = none: every compute call will fail loudly). */
int picard_device_count(void);
const char* picard_status_string(int status); /* Display text of error.rs:44-74 */
/* The library keeps the device buffers of a finished call cached for the next one (up to PICARD_CACHE_MAX_GB, default 60 % of the
 * device's memory) and one pinned host arena for a large `sources` result, because driver allocator calls are slow on these hosts
 * (an N x T buffer: 5 - 700 ms to free); this returns all of it to the driver.  The cache is also emptied when an allocation fails. */
void picard_release_cache(void);

/* ---- config (config.rs) --------------------------------------------------------------------------- */
void picard_config_default(picard_config_t* cfg);                                  /* config.rs:64-85 */
int picard_config_validate(const picard_config_t* cfg, char* err, size_t errlen);  /* config.rs:104-142 */

/* ---- fit / transform (solver.rs) ---------------------------------------------------------------- */
/* Picard::fit_with_config (solver.rs:45-189); Picard::fit (solver.rs:33) is this with the default config.
 * `x` is a HOST buffer.  With cfg->comm != NULL, `x` holds this rank's contiguous block of sample columns
 * and every rank must call collectively. */
int picard_fit(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, const picard_config_t* cfg,
               picard_result_t* out, char* err, size_t errlen);
/* Same, with `d_x` a DEVICE buffer on cfg->device (row_stride even, base 16-byte aligned).  `d_sources`
 * (nc x n_samples, leading dimension lds, may be NULL) receives the sources on the device. */
int picard_fit_device(const double* d_x, int64_t n_features, int64_t n_samples, int64_t row_stride,
                      const picard_config_t* cfg, double* d_sources, int64_t lds, picard_result_t* out, char* err,
                      size_t errlen);
/* Picard::transform (solver.rs:199-214) with PicardResult::full_unmixing (result.rs:39-44).
 * out is (n_components x n_samples), host. */
int picard_transform(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride,
                     const picard_result_t* result, double* out, int32_t device, char* err, size_t errlen);
void picard_result_free(picard_result_t* r);

/* ---- the core loop alone (core.rs:162-401), resumable; what bench.py's `value` times --------------- */
typedef struct picard_core picard_core_t;
/* d_x: preprocessed data (n x n_samples, device, this rank's shard); covariance_identity = the
 * `covariance = Some(I)` argument of core::run (solver.rs:143-147), else None. */
int picard_core_create(picard_core_t** out, const double* d_x, int64_t n, int64_t n_samples, int64_t row_stride,
                       const picard_config_t* cfg, int32_t covariance_identity, char* err, size_t errlen);
/* Runs up to `max_new_iters` further outer iterations (or to convergence / cfg->max_iter). */
int picard_core_run(picard_core_t* c, int64_t max_new_iters, int64_t* iters_done, int32_t* converged, char* err,
                    size_t errlen);
int picard_core_reset(picard_core_t* c); /* back to W = I, empty memory, iteration 0 */
/* Current state to the host: w (n x n), signs (n, may be NULL), scalars. */
int picard_core_state(picard_core_t* c, double* w, double* signs, int64_t* n_iterations, int32_t* converged,
                      double* gradient_norm, double* loss);
int picard_core_stats(picard_core_t* c, picard_stats_t* stats);
void picard_core_destroy(picard_core_t* c);

/* ---- test hooks: one evaluation point of the hot path (SURVEY.md §8a fused-pass contract) ------------ */
/* Raw moments of the pass kernels at Y = W X for host inputs (w NULL = identity):
 *   gr[i,j] = sum_t psi(y_it) y_jt ; sd[i] = sum_t psi'(y_it) ; hr[i,j] = sum_t psi'(y_it) y_jt^2 ;
 *   sq[i] = sum_t y_it^2 ; lrow[i] = sum_t loglik(y_it).        (core.rs:215-221,226,264,274 ; density.rs)
 * mode: 0 = fused (all), 1 = grad-only (lrow untouched), 2 = loss-only (gr, sd, hr untouched),
 *       3 = loss-only pass that stores Y, then the gradient moments from the stored Y (the two-kernel path an
 *           accepted loss-only line-search try takes).
 * Any output pointer may be NULL.  hr is only computed when want_h != 0. */
int picard_eval_moments(const double* x, int64_t n, int64_t n_samples, int64_t row_stride, const double* w,
                        int32_t density_kind, double alpha, int32_t mode, int32_t want_h, int32_t device, double* gr,
                        double* sd, double* hr, double* sq, double* lrow, char* err, size_t errlen);
/* picard_eval_moments with explicit execution flags (PICARD_FLAG_NO_INT8 / FORCE_INT8 ...) and the `whitened` promise of
 * core::run's covariance = Some(I); stats (may be NULL) reports which engine ran (i8_loss_passes, i8_grad_passes, i8_fallbacks). */
int picard_eval_moments_ex(const double* x, int64_t n, int64_t n_samples, int64_t row_stride, const double* w,
                           int32_t density_kind, double alpha, int32_t mode, int32_t want_h, int32_t device, uint32_t flags,
                           int32_t whitened, double* gr, double* sd, double* hr, double* sq, double* lrow, picard_stats_t* stats,
                           char* err, size_t errlen);
/* Same on a DEVICE-resident matrix, `repeats` launches timed with CUDA events on the library's stream
 * (avg_ms = mean duration of one pass incl. the partial reduction).  Bench / profiling hook. */
int picard_eval_moments_device(const double* d_x, int64_t n, int64_t n_samples, int64_t row_stride, const double* w,
                               int32_t density_kind, double alpha, int32_t mode, int32_t want_h, int32_t device,
                               int32_t repeats, double* avg_ms, double* gr, double* sd, double* hr, double* sq, double* lrow,
                               char* err, size_t errlen);
/* ... with execution flags and the `whitened` promise (bench.py's parity block compares the INT8 and FP64 engines with it). */
int picard_eval_moments_device_ex(const double* d_x, int64_t n, int64_t n_samples, int64_t row_stride, const double* w,
                                  int32_t density_kind, double alpha, int32_t mode, int32_t want_h, int32_t device, uint32_t flags,
                                  int32_t whitened, int32_t repeats, double* avg_ms, double* gr, double* sd, double* hr, double* sq,
                                  double* lrow, picard_stats_t* stats, char* err, size_t errlen);
/* The processed quantities of one iteration front (core.rs:215-293) plus the loss (core.rs:39-85) at
 * Y = W X: projected gradient g, Hessian approximation h, h_off, signs, sign_change, gradient norm, loss.
 * c = C matrix of the extended sign rule (NULL = identity); old_signs NULL = first iteration;
 * loss_signs NULL = the signs just estimated (ones when not extended). */
int picard_eval_point(const double* x, int64_t n, int64_t n_samples, int64_t row_stride, const double* w,
                      int32_t density_kind, double alpha, int32_t ortho, int32_t extended, double lambda_min,
                      const double* c, const double* old_signs, const double* loss_signs, int32_t device, double* g,
                      double* h, double* hoff, double* signs, int32_t* sign_change, double* gradient_norm, double* loss,
                      char* err, size_t errlen);

/* ---- test hooks: the N x N device kernels --------------------------------------------------------- */
int picard_matrix_exp(const double* a, int64_t n, double* out, int32_t device);        /* math.rs:38-74 */
int picard_sln_det(const double* m, int64_t n, double* sign, double* logabs, int32_t device); /* math.rs:84-88 */
int picard_sym_decorrelation(const double* w, int64_t n, double* out, int32_t device); /* math.rs:12-33 */
/* lbfgs.rs:84-133 with an explicit memory (s_list, y_list: L x n x n oldest first; r_list: L). */
int picard_compute_direction(const double* g, const double* h, const double* hoff, int64_t n, const double* s_list,
                             const double* y_list, const double* r_list, int64_t L, int32_t ortho, double* out,
                             int32_t device);
/* whitening.rs:24-35 + 48-116 on the device: mean (nf), k (nc x nf), whitened data (nc x n_samples, may be NULL). */
int picard_center_whiten(const double* x, int64_t n_features, int64_t n_samples, int64_t row_stride, int64_t n_components,
                         int32_t centering, int32_t device, double* mean, double* k, double* data, char* err,
                         size_t errlen);
/* Same for a DEVICE-resident sample matrix (this rank's column shard when comm != NULL): mean (nf, host) and
 * k (nc x nf, host) only; the caller applies them with picard_apply_device.  What bench.py uses to prepare the
 * core loop's input without a host round trip. */
int picard_center_whiten_device(const double* d_x, int64_t n_features, int64_t n_samples, int64_t row_stride,
                                int64_t n_components, int32_t centering, picard_comm_t* comm, int32_t device, double* mean,
                                double* k, char* err, size_t errlen);
/* jade.rs:22-72 on the device: x whitened (n x n_samples) host; w (n x n) out. */
int picard_jade(const double* x, int64_t n, int64_t n_samples, int64_t row_stride, int64_t max_iter, double tol,
                int32_t verbose, int32_t device, double* w, int64_t* sweeps, char* err, size_t errlen);

/* jade.rs:78-131 alone (test hook): cumulant matrices Q_ij, i <= j, as out[m][k][l] with m enumerating (i, j) row-major. */
int picard_jade_cumulants(const double* x, int64_t n, int64_t n_samples, int64_t row_stride, int32_t device, double* out, char* err,
                          size_t errlen);

/* ---- synthetic data: counter-based, identical for any shard layout ------------------ */
/* Writes sources S[i, t_offset + s] for s in [0, n_samples) into d_out (n x n_samples, leading dim ld):
 * row i < n_laplace: Laplace(b = 1/sqrt 2) (unit variance); else uniform on [-sqrt 3, sqrt 3]. */
int picard_synth_sources(double* d_out, int64_t n, int64_t n_samples, int64_t ld, int64_t t_offset, int64_t n_laplace,
                         uint64_t seed, int32_t device, void* stream);
/* d_out (n_out x n_samples, ld_out) = a (n_out x n_in, HOST, row-major) * (d_in - mean) ; mean (n_in, HOST) may be NULL. */
int picard_apply_device(const double* a, const double* mean, int64_t n_out, int64_t n_in, const double* d_in, int64_t ld_in,
                        double* d_out, int64_t ld_out, int64_t n_samples, int32_t device, void* stream);

/* Measurement hook: FP64 tensor-core (DMMA m8n8k4) peak of the device at its current clocks, in TFLOP/s, from a microbenchmark
 * of about budget_ms milliseconds.  bench.py's roofline denominator (the pool's MEASURED_PEAKS.json has no FP64 figure). */
int picard_fp64_peak_probe(int32_t device, double budget_ms, double* tflops);

/* ---- multi-GPU plumbing: one process per GPU, NCCL over NVLink ------------------------------------ */
#define PICARD_UNIQUE_ID_BYTES 128
int picard_comm_unique_id(char id[PICARD_UNIQUE_ID_BYTES]); /* call on one rank, broadcast out of band */
int picard_comm_create(picard_comm_t** out, const char id[PICARD_UNIQUE_ID_BYTES], int32_t rank, int32_t nranks,
                       int32_t device, char* err, size_t errlen);
int picard_comm_rank(const picard_comm_t* c);
int picard_comm_size(const picard_comm_t* c);
void picard_comm_destroy(picard_comm_t* c);

#ifdef __cplusplus
}
#endif
#endif /* PICARD_B200_H */
