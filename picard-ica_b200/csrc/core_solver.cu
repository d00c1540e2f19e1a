// The core Picard loop (core.rs:162-401) with device-resident state.
//
// Control flow is the reference's, quirk for quirk; what differs is where the work happens:
//   * the transform of a line-search try is folded into W: every try is ONE LOSS pass Y' = (M W) x1 that keeps Y' in HBM
//     (the "Y store", one extra N x T buffer) and returns the log-likelihood / y^2 row sums; only an ACCEPTED try is
//     followed by the stored-Y gradient pass (psi(Y') Y'^T, sum psi' [, psi'(Y') (Y'^2)^T]).  Nothing is computed for a
//     rejected point.  Without the store (no memory, PICARD_FLAG_NO_Y_STORE) the first try of an iteration runs the fused
//     from-X pass speculatively instead (pass.cuh);
//   * for whitened problems with 64 < N <= 128 both passes run on the INT8 tensor cores (i8_loss.cu, i8_grad.cu: error-free
//     digit splitting, results within ~1e-13 of the FP64 kernels -- three orders inside the parity bar, but not bit-identical
//     to them, so the iterates of the two engines agree to that level, not to the last bit) after a range check of the data;
//   * the per-row log-likelihood sums L_i are kept with each point, so the loss under new signs
//     (core.rs:317-329) and the initial loss (core.rs:185) need no extra pass.
#include "engine.cuh"
#include "i8.cuh"
#include "p2p.cuh"

#include <chrono>
#include <map>
#include <mutex>

namespace picard {

double trace_now_ms() {
  static const bool on = getenv("PICARD_TRACE") != nullptr || getenv("PICARD_TRACE_TRY") != nullptr;
  if (!on) return -1.0;
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void trace_slow(const char* what, size_t bytes, double t0_ms) {
  if (t0_ms < 0) return;
  const double dt = trace_now_ms() - t0_ms;
  if (dt > 5.0) fprintf(stderr, "[picard trace]     slow %s of %.1f MB: %.1f ms\n", what, bytes / 1048576.0, dt);
}

// PICARD_TRACE_GAPS diagnostics: host time between the moment the scalars of a kernel arrive and the moment the next kernel of the
// dependent chain has been handed to the driver (the GPU idles for at least that long), per kind of hand-over.
struct GapProbe {
  bool on = getenv("PICARD_TRACE_GAPS") != nullptr;
  double last_exit = -1.0, acc[4] = {0, 0, 0, 0}, wait_ms = 0.0;
  long n[4] = {0, 0, 0, 0}, waits = 0;
  static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  void launched(int kind) {
    if (!on || last_exit < 0) return;
    acc[kind] += now() - last_exit; ++n[kind]; last_exit = -1.0;
  }
  // device time of the N x N kernels between the passes (events on the solver's stream, read back after the next wait)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double dev_ms[2] = {0, 0};
  long dev_n[2] = {0, 0};
  int pending = -1;
  void dev_begin(int kind, cudaStream_t st) {
    if (!on) return;
    if (!ev[0]) for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[2 * kind], st);
  }
  void dev_end(int kind, cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(ev[2 * kind + 1], st);
    pending = kind;
  }
  void dev_collect() {
    if (!on || pending < 0) return;
    float ms = 0.f;
    if (cudaEventSynchronize(ev[2 * pending + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, ev[2 * pending], ev[2 * pending + 1]) == cudaSuccess) {
      dev_ms[pending] += ms; ++dev_n[pending];
    }
    pending = -1;
  }
  void report() {
    if (!on) return;
    const char* dn[2] = {"device: transforms (expm + W' products)", "device: gradient end -> front done"};
    for (int k = 0; k < 2; ++k)
      if (dev_n[k]) fprintf(stderr, "[picard gaps] %-36s %8.2f us avg over %ld\n", dn[k], 1e3 * dev_ms[k] / dev_n[k], dev_n[k]);
    const char* names[4] = {"front -> transforms enqueued", "rejected try -> next LOSS enqueued", "accepted try -> gradient enqueued", "other"};
    for (int k = 0; k < 4; ++k)
      if (n[k]) fprintf(stderr, "[picard gaps] %-36s %8.2f us avg over %ld\n", names[k], 1e3 * acc[k] / n[k], n[k]);
    if (waits) fprintf(stderr, "[picard gaps] %-36s %8.2f us avg over %ld\n", "host wait for scalars", 1e3 * wait_ms / waits, waits);
  }
};
static thread_local GapProbe g_gaps;

namespace {
struct DevCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;  // (device, capacity) -> block
  size_t cached_bytes = 0;
};
DevCache& dev_cache() { static DevCache* c = new DevCache(); return *c; }  // leaked on purpose: no destructor order issues at exit
// Cache limit per process: PICARD_CACHE_MAX_GB if set, else 60 % of the device's memory (blocks of every size are kept: returning an
// N x T buffer to the driver costs anything between 5 and 700 ms -- profiles/README.md, e2e trace of round 2).
size_t cache_limit_bytes() {
  static const size_t limit = [] {
    if (const char* e = getenv("PICARD_CACHE_MAX_GB")) return (size_t)(atof(e) * 1073741824.0);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return (size_t)4 << 30; }
    return (size_t)(0.6 * (double)total_b);
  }();
  return limit;
}
}  // namespace

void* dev_alloc(size_t bytes, size_t* capacity, int* device) {
  const size_t want = (bytes + 511) & ~(size_t)511;
  int dev = 0;
  PICARD_CUDA(cudaGetDevice(&dev));
  *device = dev;
  {
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.free_blocks.lower_bound({dev, want});
    if (it != c.free_blocks.end() && it->first.first == dev && it->first.second <= 2 * want + 4096) {
      void* p = it->second;
      *capacity = it->first.second;
      c.cached_bytes -= it->first.second;
      c.free_blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  const double t0 = trace_now_ms();
  cudaError_t e = cudaMalloc(&p, want);
  if (e == cudaErrorMemoryAllocation) {  // the cache may be what is in the way: hand it back and try once more
    cudaGetLastError();
    dev_cache_release();
    e = cudaMalloc(&p, want);
  }
  PICARD_CUDA(e);
  trace_slow("cudaMalloc", want, t0);
  *capacity = want;
  return p;
}

void dev_free(void* p, size_t capacity, int dev) {
  if (!p) return;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != dev) cudaSetDevice(dev);  // the block is filed, synchronised and freed under the device that owns it
  struct Restore { int cur, dev; ~Restore() { if (cur >= 0 && cur != dev) cudaSetDevice(cur); } } restore{cur, dev};
  if (cudaDeviceSynchronize() == cudaSuccess) {
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached_bytes + capacity <= cache_limit_bytes()) {
      c.free_blocks.insert({{dev, capacity}, p});
      c.cached_bytes += capacity;
      return;
    }
  } else {
    cudaGetLastError();
  }
  const double t0 = trace_now_ms();
  cudaFree(p);
  trace_slow("cudaFree", capacity, t0);
}

void dev_cache_release() {
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  int prev = -1;
  cudaGetDevice(&prev);
  for (auto& kv : c.free_blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
  c.free_blocks.clear();
  c.cached_bytes = 0;
  if (prev >= 0) cudaSetDevice(prev);
}

DeviceGuard::DeviceGuard(int dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(PICARD_COMPUTATION_ERROR,
                "Computation error: no usable CUDA device (libpicard_b200 has no CPU fallback): " + std::string(cudaGetErrorString(e)));
  PICARD_CUDA(cudaGetDevice(&prev));
  device = dev < 0 ? prev : dev;
  if (device >= count) throw Error(PICARD_INVALID_CONFIG, "Invalid configuration for 'device': no such CUDA device");
  PICARD_CUDA(cudaSetDevice(device));
  PICARD_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
}
DeviceGuard::~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }

CoreSolver::CoreSolver(const double* d_x, int n, int64_t t_local, int64_t ldx, const picard_config_t& cfg, bool covariance_identity,
                       int sm_count, cudaStream_t stream)
    : cfg_(cfg), d_x_(d_x), t_local_(t_local), ldx_(ldx), sm_count_(sm_count), st_(stream), comm_(cfg.comm),
      cov_identity_(covariance_identity), sc_dev_(1), sc_host_(1) {
  dims_.n = n;
  dims_.m = (int)cfg.m;
  dims_.ortho = cfg.ortho ? 1 : 0;
  dims_.extended = (cfg.extended < 0 ? cfg.ortho : cfg.extended) ? 1 : 0;
  dims_.lambda_min = cfg.lambda_min;
  // global sample count (core.rs:176 `t as f64`, with T the total over all shards)
  double tt = (double)t_local;
  if (comm_ && comm_size(comm_) > 1) {
    DevBuf<double> tmp(1);
    PICARD_CUDA(cudaMemcpyAsync(tmp.p, &tt, sizeof(double), cudaMemcpyHostToDevice, st_));
    comm_allreduce_sum(comm_, tmp.p, 1, st_);
    PICARD_CUDA(cudaMemcpyAsync(&tt, tmp.p, sizeof(double), cudaMemcpyDeviceToHost, st_));
    PICARD_CUDA(cudaStreamSynchronize(st_));
  }
  dims_.t_total = tt;
  dens_ = cfg.density_kind;
  alpha_ = (dens_ == PICARD_DENSITY_CUBE) ? 1.0 : cfg.alpha;
  need_h_ = !dims_.ortho;
  pass_padded_size(n);  // validates n

  const size_t nn = (size_t)n * n;
  const size_t msz = (size_t)mom_size(n) + MOM_EXTRA;
  size_t total = 0;
  auto take = [&](size_t count) { size_t o = total; total += (count + 1) & ~(size_t)1; return o; };
  const size_t oW = take(nn), oWt = take(nn), oM = take(nn), oD = take(nn), oC = take(nn), oG = take(nn), oGt = take(nn),
               oGo = take(nn), oH = take(nn), oho = take(n), osg = take(n), oos = take(n), oSp = take(nn), oq = take(nn),
               oms = take(nn * dims_.m), omy = take(nn * dims_.m), omr = take(2 * (size_t)dims_.m), omc = take(msz), omt = take(msz),
               olu = take(nn), oAs = take(nn), ot0 = take(nn), ot1 = take(nn), or0 = take(nn), or1 = take(nn), osl = take(small::EXPM_SLOTS);
  store_.alloc(total);
  store_.zero(st_);
  double* b = store_.p;
  W_ = b + oW; Wt_ = b + oWt; M_ = b + oM; D_ = b + oD; C_ = b + oC; G_ = b + oG; Gtmp_ = b + oGt; Gold_ = b + oGo; H_ = b + oH;
  hoff_ = b + oho; signs_ = b + osg; old_signs_ = b + oos; Sprev_ = b + oSp; q_ = b + oq; mem_s_ = b + oms; mem_y_ = b + omy;
  mem_r_ = b + omr; mom_cur_ = b + omc; mom_trial_ = b + omt; lu_work_ = b + olu;
  ew_.As = b + oAs; ew_.term0 = b + ot0; ew_.term1 = b + ot1; ew_.res0 = b + or0; ew_.res1 = b + or1; ew_.slots = b + osl;
  partial_.alloc(pass_workspace_doubles(n, sm_count_));
  if (!(cfg.flags & PICARD_FLAG_NO_Y_STORE)) {
    // one extra N x T buffer: an accepted loss-only try leaves its Y' here, so the next gradient pass skips W X.
    // Not having the memory is not an error: the gradient pass then recomputes from X.
    ldy_ = round_up(t_local_, 64);  // its own leading dimension: whole 64-sample tiles of the INT8 LOSS pass stay in bounds
    try { ybuf_.alloc((size_t)n * (size_t)ldy_); } catch (const Error&) { cudaGetLastError(); ybuf_.release(); }
  }
  PICARD_CUDA(cudaEventCreate(&ev_a_));
  PICARD_CUDA(cudaEventCreate(&ev_b_));
  PICARD_CUDA(cudaEventCreate(&ev_run0_));
  PICARD_CUDA(cudaEventCreate(&ev_run1_));
  memset(&stats_, 0, sizeof stats_);
  memset(sc_host_.p, 0, sizeof(CoreScalars));
  reset();
}

CoreSolver::~CoreSolver() {
  cudaEventDestroy(ev_a_); cudaEventDestroy(ev_b_); cudaEventDestroy(ev_run0_); cudaEventDestroy(ev_run1_);
}

void CoreSolver::reset() {
  const int n = dims_.n;
  store_.zero(st_);
  sc_dev_.zero(st_);
  stats_.kernel_launches += small::set_identity(W_, n, st_);       // core.rs:178
  stats_.kernel_launches += small::set_identity(C_, n, st_);       // core.rs:204 (replaced below when extended)
  std::vector<double> ones((size_t)n, 1.0);
  PICARD_CUDA(cudaMemcpyAsync(signs_, ones.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st_));      // core.rs:182
  PICARD_CUDA(cudaMemcpyAsync(old_signs_, ones.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st_));  // core.rs:183
  PICARD_CUDA(cudaStreamSynchronize(st_));
  iter_ = 0; n_iterations_ = 0; converged_ = false; started_ = false; have_cur_ = false; speculate_next_ = true; ybuf_valid_ = false;
  gradient_norm_ = 1.0; current_loss_ = 0.0;
}

// Decides, once per solver, whether the two passes run on the INT8 tensor cores, and prepares the digit image of x1.
//   wanted : the data is whitened (covariance = I promised by the caller) or PICARD_FLAG_FORCE_INT8 / PICARD_I8=1
//   allowed: 64 < N <= 128, no PICARD_FLAG_NO_INT8 / PICARD_I8=0, memory for the image, and -- unless forced -- the RANGE CHECK:
//            a sample is scaled by the power of two above its largest component, so a component loses the bits between that bound
//            and its own magnitude.  range = (mean bound over the samples) / (smallest row RMS of x1) is ~4 for whitened data
//            (every row has unit variance, the largest of 128 components of a sample is ~3); above 64 the FP64 kernels are used
//            (stats.i8_fallbacks = 1): the promise "whitened" is not taken on trust.
bool CoreSolver::i8_prepare() {
  if (i8_state_ != 0) return i8_state_ == 1;
  i8_state_ = -1;
  const int env = i8_env_mode();
  const bool forced = (cfg_.flags & PICARD_FLAG_FORCE_INT8) != 0 || env == 1;
  // automatic for 64 < N <= 128; PICARD_FLAG_FORCE_INT8 also admits smaller N (the kernels pad to 128 rows: measured no faster
  // than the FP64 kernels there, DESIGN.md section 7 -- kept for that measurement and for tests)
  const bool size_ok = pass_padded_size(dims_.n) == 128 || (forced && dims_.n <= 128);
  if (!size_ok || (cfg_.flags & PICARD_FLAG_NO_INT8) != 0 || env == 0) return false;
  if (!forced && !cov_identity_) return false;
  try {
    xs8_.alloc(i8_blob_bytes(t_local_));
    i8_counter_.alloc(3);
    PICARD_CUDA(cudaMemsetAsync(i8_counter_.p, 0, 3 * sizeof(unsigned int), st_));
    xstats_.alloc((size_t)I8_XSTATS);
    rowexp_.alloc(128);
  } catch (const Error&) {  // no memory for the digit image: the FP64 kernels need none
    cudaGetLastError();
    xs8_.release(); i8_counter_.release(); xstats_.release(); rowexp_.release();
    stats_.i8_fallbacks = 1;
    return false;
  }
  stats_.kernel_launches += i8_slice_x(d_x_, ldx_, t_local_, dims_.n, xs8_.p, xstats_.p, sm_count_, st_);
  std::vector<double> hs((size_t)I8_XSTATS);
  PICARD_CUDA(cudaMemcpyAsync(hs.data(), xstats_.p, sizeof(double) * I8_XSTATS, cudaMemcpyDeviceToHost, st_));
  PICARD_CUDA(cudaStreamSynchronize(st_));
  double min_ms = INFINITY;
  for (int k = 0; k < dims_.n; ++k) min_ms = std::fmin(min_ms, hs[2 + k] / (double)t_local_);
  const double mean_bound = hs[0] / (double)t_local_;
  stats_.i8_range = min_ms > 0.0 ? mean_bound / std::sqrt(min_ms) : INFINITY;
  if (!forced && !(stats_.i8_range <= 64.0)) {
    xs8_.release();
    stats_.i8_fallbacks = 1;
    return false;
  }
  i8_state_ = 1;
  return true;
}

bool CoreSolver::eval_pass(const double* d_w, int mode, bool want_h, int dens, double alpha, double* d_mom, bool store_y, int finish_which,
                           const double* finish_signs) {
  PassLaunch L;
  L.d_x = (mode == PASS_GRADY) ? ybuf_.p : d_x_; L.ldx = (mode == PASS_GRADY) ? ldy_ : ldx_; L.t_local = t_local_; L.n_in = dims_.n; L.n_out = dims_.n;
  L.d_w = d_w; L.ldw = dims_.n; L.d_bias = nullptr; L.dens = dens; L.alpha = alpha; L.mode = mode; L.want_h = want_h;
  L.d_partial = partial_.p; L.d_mom = d_mom; L.sm_count = sm_count_; L.stream = st_;
  const bool store = store_y && mode == PASS_LOSS && ybuf_.p != nullptr;
  L.d_out = store ? ybuf_.p : nullptr; L.ld_out = store ? ldy_ : 0;
  if (mode == PASS_LOSS) ybuf_valid_ = store;
  // INT8 tensor-core engines (i8_loss.cu / i8_grad.cu) for whitened problems with 64 < N <= 128
  const bool use_i8 = mode == PASS_LOSS && dens != DENS_LINEAR && i8_prepare();
  const bool use_i8_grad = mode == PASS_GRADY && i8_state_ == 1 && i8_grad_supported(dims_.n, dens, want_h);
  if (use_i8_grad) stats_.kernel_launches += i8_row_exponents(d_w, dims_.n, xstats_.p, rowexp_.p, st_);
  if (mode == PASS_FUSED) stats_.fused_passes++;
  else if (mode == PASS_GRAD) stats_.grad_passes++;
  else if (mode == PASS_GRADY) stats_.grady_passes++;
  else stats_.loss_passes++;
  resolve_pass_time();
  PICARD_CUDA(cudaEventRecord(ev_a_, st_));
  bool finished = false, exchanged = false;
  const bool single = !(comm_ && comm_size(comm_) > 1);
  P2PCall px;
  if (use_i8) {
    // one launch: W' digits, streaming, deterministic reduction of the partials, [exchange between the ranks over peer memory,]
    // loss + accept flag + publish -- on any number of GPUs when the peer mailboxes are available
    I8LossFinish fin;
    fin.counter = i8_counter_.p; fin.counter_total = &i8_counter_total_[0];
    if (!single && comm_p2p_next(comm_, 2 * (size_t)dims_.n, &px)) { fin.px = &px; exchanged = true; }
    if (finish_which >= 0 && (single || exchanged)) {
      fin.finish = 1; fin.which = finish_which; fin.dims = &dims_; fin.signs = finish_signs; fin.sc = sc_dev_.p; fin.sc_map = sc_host_.p;
      fin.seq = next_seq();
      finished = true;
    }
    stats_.kernel_launches += launch_loss_i8(L, xs8_.p, fin);
    stats_.i8_loss_passes++;
  } else if (use_i8_grad) {
    const size_t nn = (size_t)dims_.n * dims_.n;
    if (!single && 2 * (L.sm_count / 2) <= L.sm_count && comm_p2p_next(comm_, nn + (size_t)dims_.n, &px)) exchanged = true;
    stats_.kernel_launches += launch_grad_i8(L, rowexp_.p, i8_counter_.p + 1, &i8_counter_total_[1], exchanged ? &px : nullptr);
    stats_.i8_grad_passes++;
  }
  else stats_.kernel_launches += launch_pass(L);
  g_gaps.launched(mode == PASS_LOSS ? 1 : (mode == PASS_GRADY ? 2 : 3));
  PICARD_CUDA(cudaEventRecord(ev_b_, st_));
  // one allreduce of exactly what this pass produced -- unless the pass kernel has carried the exchange itself
  if (comm_ && comm_size(comm_) > 1 && !exchanged) {
    const int n = dims_.n;
    const size_t nn = (size_t)n * n;
    if (mode == PASS_LOSS) comm_allreduce_sum(comm_, d_mom + mom_off_sq(n), 2 * (size_t)n, st_);
    else if (mode == PASS_FUSED) comm_allreduce_sum(comm_, d_mom, nn + 3 * (size_t)n + (want_h ? nn : 0), st_);
    // PASS_GRAD and PASS_GRADY produce [Gr, Sd, Sq] (+ Hr)
    else comm_allreduce_sum2(comm_, d_mom, nn + 2 * (size_t)n, want_h ? d_mom + mom_off_hr(n) : nullptr, want_h ? nn : 0, st_);
  }
  last_pass_mode_ = mode;
  return finished;
}

bool CoreSolver::pass(const double* d_w, int mode, double* d_mom, int finish_which, const double* finish_signs) {
  return eval_pass(d_w, mode, need_h_, dens_, alpha_, d_mom, /*store_y=*/true, finish_which, finish_signs);  // LOSS mode: want_h = the Sq row sums only
}

// The kernel that finalised the scalars (loss_kernel / front_kernel, launched with next_seq()) has written them into the pinned
// host mirror and then the sequence number: spin on it instead of a D2H copy + stream synchronisation.
void CoreSolver::fetch_scalars() {
  volatile CoreScalars* h = sc_host_.p;
  const double g0 = g_gaps.on ? GapProbe::now() : 0.0;
  struct GapExit { double g0; ~GapExit() { if (g_gaps.on) { g_gaps.dev_collect(); g_gaps.last_exit = GapProbe::now(); g_gaps.wait_ms += g_gaps.last_exit - g0; ++g_gaps.waits; } } } gap_exit{g0};
  for (uint64_t spins = 1; h->seq != seq_; ++spins) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0x3FFF) == 0) {  // a faulting kernel must not leave the host spinning forever
      const cudaError_t e = cudaStreamQuery(st_);
      if (e != cudaSuccess && e != cudaErrorNotReady)
        throw Error(PICARD_COMPUTATION_ERROR, std::string("Computation error: CUDA failure '") + cudaGetErrorString(e) + "' in the core loop");
      if (e == cudaSuccess && h->seq != seq_) {  // stream idle: the publishing store is visible by now (system-scope fence)
        PICARD_CUDA(cudaStreamSynchronize(st_));
        if (h->seq != seq_) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the core loop's scalars were not published");
      }
    }
  }
}

// Device time of the last pass kernel (events around its launch) goes to the statistics.  The fused LOSS kernel publishes its
// scalars before it retires, so the closing event may still be pending when the host has its answer: the time is resolved lazily,
// at the next pass launch or when the loop returns (by then the event has completed; no extra wait on the critical path).
void CoreSolver::resolve_pass_time() {
  if (last_pass_mode_ < 0) return;
  float ms = 0.f;
  if (cudaEventSynchronize(ev_b_) == cudaSuccess && cudaEventElapsedTime(&ms, ev_a_, ev_b_) == cudaSuccess) {
    if (last_pass_mode_ == PASS_FUSED) stats_.pass_ms_fused += ms;
    else if (last_pass_mode_ == PASS_GRAD) stats_.pass_ms_grad += ms;
    else if (last_pass_mode_ == PASS_LOSS) stats_.pass_ms_loss += ms;
    else if (last_pass_mode_ == PASS_GRADY) stats_.pass_ms_grady += ms;
  } else {
    cudaGetLastError();
  }
  last_pass_mode_ = -1;
}

// One line-search try (core.rs:118-128): transform, W' = M W, pass at W', loss, accept flag -> scalars on host.
void CoreSolver::try_point(double alpha, bool speculate, int try_index, int tries_planned) {
  const int n = dims_.n;
  stats_.ls_tries++;
  static const bool prof = getenv("PICARD_TRACE_TRY") != nullptr;  // diagnostics: where a line-search try spends its time
  static cudaEvent_t pe[4];
  static bool pe_init = false;
  static double acc_ms[4] = {0, 0, 0, 0};
  static long acc_n = 0;
  const double w0 = prof ? trace_now_ms() : 0.0;
  if (prof && !pe_init) { for (auto& e : pe) cudaEventCreate(&e); pe_init = true; }
  if (prof) cudaEventRecord(pe[0], st_);
  w_try_ = Wt_;
  if (dims_.ortho) {  // W' = expm(alpha D) W (core.rs:119,125)
    if (try_index == 0) {  // every candidate alpha / 2^t of this search from one Taylor run (bit-identical to one run per try)
      if (!wt_all_.p) wt_all_.alloc((size_t)small::EXPM_NC * n * n);
      // the first candidates only: 95 % of the line searches end within four tries (2.4 on average at c3), and every candidate costs
      // one more N x N x N product in the kernel; later tries take the per-try kernel (same bits)
      g_gaps.dev_begin(0, st_);
      const int r = small::matrix_exp_candidates(D_, alpha, sc_host_.p->norm_d, n, tries_planned < 4 ? tries_planned : 4, ew_, W_, wt_all_.p, st_);
      g_gaps.dev_end(0, st_);
      cand_ready_ = r > 0 ? r : 0;
      if (r > 0) stats_.kernel_launches += 1;
      g_gaps.launched(0);
    }
    if (try_index < cand_ready_) w_try_ = wt_all_.p + (size_t)try_index * n * n;
    else stats_.kernel_launches += small::matrix_exp(D_, alpha, sc_host_.p->norm_d, n, ew_, nullptr, st_, W_, Wt_);
  } else {
    stats_.kernel_launches += small::eye_plus_scaled(D_, alpha, M_, n, st_);                      // core.rs:121
    stats_.kernel_launches += small::matmul(M_, W_, Wt_, n, false, 1.0, false, st_);              // core.rs:125
  }
  if (!dims_.ortho)  // -log|det W'| term of the loss (core.rs:51-70)
    stats_.kernel_launches += small::sln_det(Wt_, n, lu_work_, mom_trial_ + mom_size(n), st_);
  static double host_enq_ms = 0.0;
  if (prof) { host_enq_ms += trace_now_ms() - w0; cudaEventRecord(pe[1], st_); }
  const bool finished = pass(w_try_, speculate ? PASS_FUSED : PASS_LOSS, mom_trial_, 0, signs_);  // core.rs:124,127
  if (prof) cudaEventRecord(pe[2], st_);
  if (!finished) stats_.kernel_launches += small::loss_from_moments(dims_, mom_trial_, signs_, sc_dev_.p, 0, st_, sc_host_.p, next_seq());
  if (prof) cudaEventRecord(pe[3], st_);
  fetch_scalars();
  if (prof) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, pe[0], pe[1]); cudaEventElapsedTime(&b, pe[1], pe[2]); cudaEventElapsedTime(&c, pe[2], pe[3]);
    acc_ms[0] += a; acc_ms[1] += b; acc_ms[2] += c; acc_ms[3] += trace_now_ms() - w0; ++acc_n;
    if (acc_n % 16 == 0)
      fprintf(stderr, "[picard trace] try: transform %.3f ms (host enqueue %.3f ms), pass+reduce(+allreduce) %.3f ms, loss kernel %.3f ms, host wall %.3f ms (avg of %ld)\n",
              acc_ms[0] / acc_n, host_enq_ms / acc_n, acc_ms[1] / acc_n, acc_ms[2] / acc_n, acc_ms[3] / acc_n, acc_n);
  }
}

int64_t CoreSolver::run(int64_t max_new) {
  const int n = dims_.n;
  // Line-search policy.  With the Y store every try is a LOSS pass that keeps Y' (2 N^2 T flop) and only an ACCEPTED try
  // pays for the gradient moments (stored-Y pass, 2 N^2 T): nothing is ever computed for a rejected point.  Without the
  // store (no memory / flag) the first try of an iteration speculatively runs the FUSED pass instead.
  const bool can_fuse = pass_padded_size(n) <= 128;
  const bool force_spec = (cfg_.flags & PICARD_FLAG_FORCE_SPECULATION) != 0 && can_fuse;
  const bool no_spec = (cfg_.flags & PICARD_FLAG_NO_SPECULATION) != 0 || !can_fuse || (ybuf_.p != nullptr && !force_spec);
  PICARD_CUDA(cudaEventRecord(ev_run0_, st_));
  if (!started_) {
    // initial loss with signs = 1 (core.rs:185-194, quirk Q1); the same pass already yields the first gradient
    if (!dims_.ortho) stats_.kernel_launches += small::sln_det(W_, n, lu_work_, mom_cur_ + mom_size(n), st_);
    const bool finished = pass(W_, no_spec ? PASS_LOSS : PASS_FUSED, mom_cur_, 1, nullptr);
    have_cur_ = !no_spec;
    if (!finished) stats_.kernel_launches += small::loss_from_moments(dims_, mom_cur_, nullptr, sc_dev_.p, 1, st_, sc_host_.p, next_seq());
    fetch_scalars();
    if (sc_host_.p->loss_singular) throw Error(PICARD_SINGULAR_MATRIX, "Singular matrix encountered during computation");
    current_loss_ = sc_host_.p->current_loss;
    if (dims_.extended && !cov_identity_) {
      // C = Y Y^T / T once, never updated (core.rs:199-205, quirk Q5): a LINEAR gradient pass gives sum y y^T
      eval_pass(W_, PASS_GRAD, false, DENS_LINEAR, 1.0, mom_trial_);
      stats_.kernel_launches += small::copy_scaled(mom_trial_ + mom_off_gr(n), C_, (int64_t)n * n, 1.0 / dims_.t_total, st_);
    }
    started_ = true;
  }
  int64_t done = 0;
  while (done < max_new && iter_ < cfg_.max_iter && !converged_) {
    if (!have_cur_) {  // gradient moments of the current iterate are missing (a loss-only try was accepted)
      pass(W_, ybuf_valid_ ? PASS_GRADY : PASS_GRAD, mom_cur_);  // its Y' is still in ybuf_: no W X product needed
      have_cur_ = true;
    }
    small::FrontArgs fa;
    fa.d = dims_; fa.mom = mom_cur_; fa.C = C_; fa.G = G_; fa.Gtmp = Gtmp_; fa.G_old = Gold_; fa.H = H_; fa.hoff = hoff_;
    fa.signs = signs_; fa.old_signs = old_signs_; fa.S_prev = Sprev_; fa.mem_s = mem_s_; fa.mem_y = mem_y_; fa.mem_r = mem_r_;
    fa.q = q_; fa.D = D_; fa.sc = sc_dev_.p; fa.first_iter = (iter_ == 0) ? 1 : 0; fa.do_lbfgs = 1;
    fa.sc_map = sc_host_.p; fa.seq = next_seq();
    g_gaps.dev_begin(1, st_);
    stats_.kernel_launches += small::iteration_front(fa, st_);
    g_gaps.dev_end(1, st_);
    fetch_scalars();
    gradient_norm_ = sc_host_.p->gradient_norm;
    if (sc_host_.p->sign_change) stats_.sign_changes++;
    n_iterations_ = iter_ + 1;                              // core.rs:212,396 (quirk Q10)
    if (gradient_norm_ < cfg_.tol) { converged_ = true; break; }  // core.rs:289-293
    current_loss_ = sc_host_.p->current_loss;

    // ---- backtracking line search (core.rs:99-150)
    double alpha = 1.0;
    bool success = false;
    bool last_spec = false;
    bool first_try_ok = false;
    for (int64_t t = 0; t < cfg_.ls_tries; ++t) {
      last_spec = !no_spec && t == 0 && speculate_next_;
      try_point(alpha, last_spec, (int)t, (int)cfg_.ls_tries);
      if (sc_host_.p->accept) { success = true; first_try_ok = (t == 0); break; }
      alpha /= 2.0;
    }
    if (!success) {  // gradient-descent fallback (core.rs:349-367, quirk Q3): 10 tries, result taken regardless
      stats_.fallbacks++;
      stats_.kernel_launches += small::negate_into(G_, D_, (int64_t)n * n, sc_dev_.p, st_);  // also clears the memory
      sc_host_.p->norm_d = gradient_norm_;
      alpha = 1.0;
      for (int t = 0; t < 10; ++t) {
        last_spec = false;
        try_point(alpha, false, t, 10);
        if (sc_host_.p->accept) { success = true; break; }
        alpha /= 2.0;
      }
    }
    // step = direction * alpha (on failure alpha has been halved once more: quirk Q2)
    stats_.kernel_launches += small::accept_step(dims_, D_, alpha, Sprev_, w_try_, C_, (dims_.extended && cov_identity_) ? 1 : 0,
                                                 sc_dev_.p, st_);
    if (w_try_ == Wt_) std::swap(W_, Wt_);
    else PICARD_CUDA(cudaMemcpyAsync(W_, w_try_, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, st_));
    std::swap(mom_cur_, mom_trial_);
    have_cur_ = last_spec;
    speculate_next_ = first_try_ok;
    current_loss_ = sc_host_.p->new_loss;
    if (cfg_.verbose) {
      auto e4 = [](double v) {
        char buf[64]; snprintf(buf, sizeof buf, "%.4e", v);
        std::string s(buf); size_t e = s.find('e');
        if (e == std::string::npos) return s;
        return s.substr(0, e) + "e" + std::to_string(atoi(s.c_str() + e + 1));
      };
      if (!comm_ || comm_rank(comm_) == 0) {
        printf("iteration %lld, gradient norm = %s, loss = %s\n", (long long)(iter_ + 1), e4(gradient_norm_).c_str(),
               e4(current_loss_).c_str());
        fflush(stdout);
      }
    }
    ++iter_;
    ++done;
  }
  PICARD_CUDA(cudaEventRecord(ev_run1_, st_));
  PICARD_CUDA(cudaStreamSynchronize(st_));
  resolve_pass_time();
  float ms = 0.f;
  PICARD_CUDA(cudaEventElapsedTime(&ms, ev_run0_, ev_run1_));
  stats_.core_ms += ms;
  g_gaps.report();
  return done;
}

// ica_par (solver.rs:218-249): W <- symdecor(w_init); repeat: C = E[g(WX)X^T] - diag(E[g'(WX)]) W ; W <- symdecor(C).
// E[g(WX) X^T] is the Gr moment of the ordinary gradient pass times W (W is orthogonal after symdecor, so
// X^T = Y^T W^-T = Y^T W): no extra kernel for the N x T part.
void CoreSolver::fastica(int64_t iters, double* w_host) {
  const int n = dims_.n;
  const size_t nn = (size_t)n * n;
  DevBuf<double> work(small::sym_decorrelation_work(n)), tmp(nn);
  DevBuf<int> d_status(1);
  auto decorrelate = [&](const double* src, double* dst) {
    stats_.kernel_launches += small::sym_decorrelation(src, n, work.p, dst, d_status.p, st_);
    int status = 0;
    PICARD_CUDA(cudaMemcpyAsync(&status, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
    PICARD_CUDA(cudaStreamSynchronize(st_));
    if (status != PICARD_OK) throw Error(PICARD_SINGULAR_MATRIX, "Singular matrix encountered during computation");  // math.rs:21-24
  };
  PICARD_CUDA(cudaMemcpyAsync(Wt_, w_host, sizeof(double) * nn, cudaMemcpyHostToDevice, st_));
  decorrelate(Wt_, W_);
  const bool from_x = pass_padded_size(n) <= 128;
  for (int64_t it = 0; it < iters; ++it) {
    if (from_x) {
      eval_pass(W_, PASS_GRAD, false, dens_, alpha_, mom_cur_);
    } else {  // N > 128: Y = W X kept by a LOSS pass, moments from the stored Y
      if (!ybuf_.p) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: no memory for the Y store (needed for N > 128)");
      eval_pass(W_, PASS_LOSS, false, dens_, alpha_, mom_cur_, true);
      eval_pass(W_, PASS_GRADY, false, dens_, alpha_, mom_cur_);
    }
    stats_.kernel_launches += small::fastica_matrix(mom_cur_, n, dims_.t_total, W_, tmp.p, Wt_, st_);
    decorrelate(Wt_, W_);
  }
  PICARD_CUDA(cudaMemcpyAsync(w_host, W_, sizeof(double) * nn, cudaMemcpyDeviceToHost, st_));
  PICARD_CUDA(cudaStreamSynchronize(st_));
}

void CoreSolver::hook_moments(const double* w_host, int mode, bool want_h, double* gr, double* sd, double* hr, double* sq,
                              double* lrow) {
  const int n = dims_.n;
  const size_t nn = (size_t)n * n;
  if (w_host) PICARD_CUDA(cudaMemcpyAsync(W_, w_host, sizeof(double) * nn, cudaMemcpyHostToDevice, st_));
  if (mode == 3) {  // the two-kernel path of an accepted loss-only try: LOSS pass storing Y', then moments from the stored Y'
    if (!ybuf_.p) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: no memory for the Y store");
    eval_pass(W_, PASS_LOSS, want_h, dens_, alpha_, mom_cur_, true);
    eval_pass(W_, PASS_GRADY, want_h, dens_, alpha_, mom_cur_);
    mode = PASS_FUSED;  // all sections are valid now
  } else {
    eval_pass(W_, mode, want_h, dens_, alpha_, mom_cur_);
  }
  auto get = [&](double* dst, int64_t off, size_t cnt) {
    if (dst) PICARD_CUDA(cudaMemcpyAsync(dst, mom_cur_ + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st_));
  };
  if (mode != PASS_LOSS) { get(gr, mom_off_gr(n), nn); get(sd, mom_off_sd(n), n); if (want_h) get(hr, mom_off_hr(n), nn); }
  get(sq, mom_off_sq(n), n);
  if (mode != PASS_GRAD) get(lrow, mom_off_ll(n), n);
  PICARD_CUDA(cudaStreamSynchronize(st_));
}

void CoreSolver::hook_point(const double* w_host, const double* c_host, const double* old_signs_host, const double* loss_signs_host,
                            double* g, double* h, double* hoff, double* signs, int32_t* sign_change, double* gradient_norm,
                            double* loss) {
  const int n = dims_.n;
  const size_t nn = (size_t)n * n;
  if (w_host) PICARD_CUDA(cudaMemcpyAsync(W_, w_host, sizeof(double) * nn, cudaMemcpyHostToDevice, st_));
  if (c_host) PICARD_CUDA(cudaMemcpyAsync(C_, c_host, sizeof(double) * nn, cudaMemcpyHostToDevice, st_));
  if (old_signs_host) PICARD_CUDA(cudaMemcpyAsync(old_signs_, old_signs_host, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
  if (!dims_.ortho) stats_.kernel_launches += small::sln_det(W_, n, lu_work_, mom_cur_ + mom_size(n), st_);
  if (ybuf_.p) {  // the product path: LOSS pass keeping Y', then the stored-Y gradient pass
    eval_pass(W_, PASS_LOSS, need_h_, dens_, alpha_, mom_cur_, true);
    eval_pass(W_, PASS_GRADY, need_h_, dens_, alpha_, mom_cur_);
  } else {
    eval_pass(W_, PASS_FUSED, need_h_, dens_, alpha_, mom_cur_);
  }
  small::FrontArgs fa;
  fa.d = dims_; fa.mom = mom_cur_; fa.C = C_; fa.G = G_; fa.Gtmp = Gtmp_; fa.G_old = Gold_; fa.H = H_; fa.hoff = hoff_;
  fa.signs = signs_; fa.old_signs = old_signs_; fa.S_prev = Sprev_; fa.mem_s = mem_s_; fa.mem_y = mem_y_; fa.mem_r = mem_r_;
  fa.q = q_; fa.D = D_; fa.sc = sc_dev_.p; fa.first_iter = old_signs_host ? 0 : 1; fa.do_lbfgs = 0;
  stats_.kernel_launches += small::iteration_front(fa, st_);
  const double* ls = signs_;
  if (loss_signs_host) {
    PICARD_CUDA(cudaMemcpyAsync(q_, loss_signs_host, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
    ls = q_;
  }
  stats_.kernel_launches += small::loss_from_moments(dims_, mom_cur_, ls, sc_dev_.p, 1, st_, sc_host_.p, next_seq());
  fetch_scalars();
  auto get = [&](double* dst, const double* src, size_t cnt) {
    if (dst) PICARD_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st_));
  };
  get(g, G_, nn); get(h, H_, nn); get(hoff, hoff_, n); get(signs, signs_, n);
  PICARD_CUDA(cudaStreamSynchronize(st_));
  if (sign_change) *sign_change = sc_host_.p->sign_change;
  if (gradient_norm) *gradient_norm = sc_host_.p->gradient_norm;
  if (loss) *loss = sc_host_.p->current_loss;
}

void CoreSolver::state(double* w, double* signs, int64_t* n_iterations, int32_t* converged, double* gradient_norm, double* loss) {
  const int n = dims_.n;
  if (w) PICARD_CUDA(cudaMemcpyAsync(w, W_, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st_));
  if (signs) PICARD_CUDA(cudaMemcpyAsync(signs, signs_, sizeof(double) * n, cudaMemcpyDeviceToHost, st_));
  PICARD_CUDA(cudaStreamSynchronize(st_));
  if (n_iterations) *n_iterations = n_iterations_;
  if (converged) *converged = converged_ ? 1 : 0;
  if (gradient_norm) *gradient_norm = gradient_norm_;
  if (loss) *loss = current_loss_;
}

}  // namespace picard
