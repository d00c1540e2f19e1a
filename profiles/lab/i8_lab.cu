// profiles/lab/i8_lab.cu -- measurement harness for the two INT8 tensor-core kernels of the library (NOT part of the product:
// it includes the product's kernel headers and instantiates ablation / trace variants the library does not ship).
//   correctness : Y' of loss_i8_kernel and Gr of grad_i8_kernel (both operand layouts) against naive f64 device kernels
//   timing      : CUDA events, `reps` launches after 2 warm-ups, kernels alone (no reduction, no slicing of W')
//   ablations   : loss: 1 = no density, 2 = no Y' store, 3 = neither; grad: 1 = no psi evaluation
//   trace       : clock64() at the phase boundaries of CTA 0 (MMA thread and one converter / epilogue warp)
// Build: make -C profiles/lab.   Run: profiles/lab/i8_lab [T] [reps]  (one JSON object per line on stdout)
#define I8_TRACE_SLOTS 96
#include <cuda.h>

#include <cstdio>
#include <vector>

#include "../../picard-ica_b200/csrc/i8_grad_kernel.cuh"
#include "../../picard-ica_b200/csrc/i8_loss_kernel.cuh"

using namespace picard;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"cuda_error\": \"%s\", \"at\": \"%s\"}\n", cudaGetErrorString(e_), #x); exit(1); } } while (0)

namespace picard {
CUtensorMap make_tmap_box(const double* d_x, int64_t ldx, int64_t t_local, int n_in, int box_cols, int np, bool swizzle128) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)t_local, (cuuint64_t)n_in};
  cuuint64_t gstr[1] = {(cuuint64_t)ldx * 8};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)np};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((Fn)sym)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(d_x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("{\"tmap_error\": %d}\n", (int)r); exit(1); }
  return m;
}
CUtensorMap make_tmap(const double* d_x, int64_t ldx, int64_t t_local, int n_in, int np) { return make_tmap_box(d_x, ldx, t_local, n_in, 16, np, true); }
}  // namespace picard

__device__ __forceinline__ double hash_normal(uint64_t k) {  // sum of 4 uniforms, unit variance
  double s = 0;
  for (int i = 0; i < 4; ++i) {
    k += 0x9E3779B97F4A7C15ull; uint64_t z = k; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    s += (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
  return s * 1.7320508075688772;
}
__global__ void fill_kernel(double* x, int n, int64_t t, int64_t ld, uint64_t seed, double scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)n * t; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / t); const int64_t c = i % t;
    double v = hash_normal(seed + (uint64_t)i * 4);
    if (r & 1) v = v * v * v * 0.4;  // heavier tails on odd rows
    x[(size_t)r * ld + c] = v * scale;
  }
}
__global__ void naive_y_kernel(const double* w, const double* x, int n, int64_t ld, int64_t t_check, double* y) {  // y (n x t_check)
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * t_check) return;
  const int r = (int)(i / t_check); const int64_t c = i % t_check;
  double s = 0;
  for (int k = 0; k < n; ++k) s = fma(w[r * n + k], x[(size_t)k * ld + c], s);
  y[i] = s;
}
__global__ void naive_g_kernel(const double* y, int n, int64_t ld, int64_t t, double* g, double* sd) {  // g[i][j] = sum_t tanh(y_it) y_jt
  const int i = blockIdx.x, j = threadIdx.x;
  double s = 0, d = 0;
  for (int64_t c = 0; c < t; ++c) { const double p = tanh(y[(size_t)i * ld + c]); s = fma(p, y[(size_t)j * ld + c], s); d += 1.0 - p * p; }
  g[i * n + j] = s;
  if (j == 0) sd[i] = d;
}

// full-size comparison of two n x t matrices: number of elements that differ by more than tol, the largest difference, the first bad index
__global__ void compare_kernel(const double* a, int64_t lda, const double* b, int64_t ldb, int n, int64_t t, double tol, unsigned long long* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)n * t; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / t); const int64_t c = i % t;
    const double d = fabs(a[(size_t)r * lda + c] - b[(size_t)r * ldb + c]);
    if (!(d <= tol)) {
      atomicAdd(&out[0], 1ull);
      atomicMax(&out[1], (unsigned long long)__double_as_longlong(d));
      atomicMin(&out[2], (unsigned long long)i);
    }
  }
}
__global__ void checksum_kernel(const double* a, int64_t lda, int n, int64_t t, unsigned long long* out) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)n * t; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / t); const int64_t c = i % t;
    acc += (unsigned long long)__double_as_longlong(a[(size_t)r * lda + c]) * (2ull * (unsigned long long)i + 1ull);
  }
  atomicAdd(out, acc);
}
template <typename F>
static float time_ms(F&& launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

static void print_trace(const char* what, const std::vector<long long>& tr, const int* slots, const char* const* names, int nslots, int tiles) {
  // mean difference between consecutive phase stamps over tiles [8, tiles), and the tile period
  printf("{\"trace\": \"%s\"", what);
  for (int s = 0; s + 1 < nslots; ++s) {
    double acc = 0; int cnt = 0;
    for (int it = 8; it < tiles; ++it) { acc += (double)(tr[it * 8 + slots[s + 1]] - tr[it * 8 + slots[s]]); ++cnt; }
    printf(", \"%s\": %.0f", names[s], acc / cnt);
  }
  double per = (double)(tr[(tiles - 1) * 8 + slots[0]] - tr[8 * 8 + slots[0]]) / (tiles - 1 - 8);
  printf(", \"tile_period_cycles\": %.0f}\n", per);
}

template <int ABL>
static void run_loss(const uint8_t* xblob, const double* w, double* yout, int64_t ld, int64_t t, int n, double* partial, int sms, int reps,
                     long long* d_trace) {
  using G = i8::LossGeom<I8_TILE>;
  auto kern = i8::loss_i8_kernel<DENS_TANH, false, I8_TILE, ABL>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
  const CUtensorMap tm = make_tmap_box(yout, ld, t, n, G::CPT, 32, true);
  PassParams p{};
  p.w = w; p.ldw = n; p.n_out = n; p.n_in = n; p.t_local = t; p.n_tiles = (t + I8_TILE - 1) / I8_TILE; p.dp = make_dens_params(DENS_TANH, 1.0); p.partial = partial; p.out = yout; p.ld_out = ld;
  const int grid = sms;
  float ms = time_ms([&] { kern<<<grid, G::NTHREADS, G::SMEM_BYTES>>>(xblob, tm, p, i8::LossTail{}, d_trace); }, reps);
  CK(cudaGetLastError());
  printf("{\"kernel\": \"loss_i8\", \"ablation\": %d, \"T\": %lld, \"ms\": %.4f, \"ms_at_1e7\": %.3f}\n", ABL, (long long)t, ms, ms * 1e7 / (double)t);
  if (ABL & 4) {
    std::vector<long long> tr(I8_TRACE_SLOTS * 8);
    CK(cudaMemcpy(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
    const int s1[] = {0, 1, 2, 3}; const char* const n1[] = {"wait_tile", "wait_acc_empty", "issue"};
    print_trace("loss: MMA thread", tr, s1, n1, 4, I8_TRACE_SLOTS);
    const int s2[] = {4, 5, 6, 7}; const char* const n2[] = {"wait_acc_full", "tmem_ld", "math+store"};
    print_trace("loss: epilogue warp 0", tr, s2, n2, 4, I8_TRACE_SLOTS);
  }
}

template <int ABL, int LAYOUT>
static float run_grad(const double* y, int64_t ld, int64_t t, int n, const int* rowexp, double* partial, int sms, int reps, long long* d_trace) {
  using G = i8::GradGeom;
  auto kern = i8::grad_i8_kernel<DENS_TANH, ABL, LAYOUT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
  const CUtensorMap tm = make_tmap(y, ld, t, n, 128);
  i8::GradParams p{};
  p.n = n; p.t_local = t; p.n_tiles = (t + 31) / 32; p.dp = make_dens_params(DENS_TANH, 1.0); p.rowexp = rowexp; p.psi_exp = i8::psi_exponent(DENS_TANH, 1.0);
  p.partial = partial; p.counter = nullptr; p.target = 0; p.mom = nullptr;
  int64_t n_tg = sms / 2; if (n_tg > p.n_tiles) n_tg = p.n_tiles;
  float ms = time_ms([&] { kern<<<(unsigned)(2 * n_tg), G::NTHREADS, G::SMEM_BYTES>>>(tm, p, d_trace); }, reps);
  CK(cudaGetLastError());
  printf("{\"kernel\": \"grad_i8\", \"ablation\": %d, \"layout\": %d, \"T\": %lld, \"ms\": %.4f, \"ms_at_1e7\": %.3f}\n", ABL, LAYOUT, (long long)t, ms,
         ms * 1e7 / (double)t);
  if (ABL & 4) {
    std::vector<long long> tr(I8_TRACE_SLOTS * 8);
    CK(cudaMemcpy(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
    const int s1[] = {0, 1, 2}; const char* const n1[] = {"wait_digits", "issue"};
    print_trace("grad: MMA thread", tr, s1, n1, 3, I8_TRACE_SLOTS);
    const int s2[] = {4, 5, 6, 7}; const char* const n2[] = {"wait_y", "convert(+wait slot)", "store+signal"};
    print_trace("grad: converter warp 0", tr, s2, n2, 4, I8_TRACE_SLOTS);
  }
  return ms;
}

// sum the per-CTA partials of the gradient kernel on the host: G[i][j], Sd[i]
static void reduce_grad(const std::vector<double>& part, int n_tg, int n, std::vector<double>& g, std::vector<double>& sd) {
  const int psz = 64 * 128 + 3 * 64;
  g.assign((size_t)n * n, 0.0); sd.assign(n, 0.0);
  for (int k = 0; k < n_tg; ++k)
    for (int h = 0; h < 2; ++h) {
      const double* pp = part.data() + (size_t)(2 * k + h) * psz;
      for (int il = 0; il < 64; ++il) {
        const int i = 64 * h + il;
        if (i >= n) continue;
        for (int j = 0; j < n; ++j) g[(size_t)i * n + j] += pp[il * 128 + j];
        sd[i] += pp[64 * 128 + il];
      }
    }
}

int main(int argc, char** argv) {
  const int64_t T = argc > 1 ? (int64_t)atof(argv[1]) : 2000000;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const int only = argc > 3 ? atoi(argv[3]) : 0;  // 1: one LOSS launch + one gradient launch only (for ncu captures)
  const int n = 128;
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ld = (T + 63) / 64 * 64;
  double *x, *y, *w, *partial, *xstats, *ynaive, *gnaive, *sdnaive;
  uint8_t* xblob;
  int* rowexp;
  long long* d_trace;
  CK(cudaMalloc(&x, sizeof(double) * n * ld)); CK(cudaMalloc(&y, sizeof(double) * n * ld));
  CK(cudaMemset(y, 0, sizeof(double) * n * ld));
  CK(cudaMalloc(&w, sizeof(double) * n * n));
  CK(cudaMalloc(&partial, sizeof(double) * sms * 2 * (2 * 128 * 128 + 3 * 128)));
  CK(cudaMalloc(&xstats, sizeof(double) * I8_XSTATS));
  CK(cudaMalloc(&xblob, (size_t)((T + I8_TILE - 1) / I8_TILE) * i8::LossGeom<I8_TILE>::TILE_BYTES + 1024));
  CK(cudaMalloc(&rowexp, 128 * sizeof(int)));
  CK(cudaMalloc(&d_trace, I8_TRACE_SLOTS * 8 * sizeof(long long))); CK(cudaMemset(d_trace, 0, I8_TRACE_SLOTS * 8 * sizeof(long long)));
  fill_kernel<<<sms * 8, 256>>>(x, n, T, ld, 1234, 1.0);
  fill_kernel<<<64, 256>>>(w, n, n, n, 99, 1.0 / sqrt((double)n));
  CK(cudaDeviceSynchronize());

  // ---- slicing
  const int64_t n_tiles = (T + I8_TILE - 1) / I8_TILE;
  CK(cudaMemset(xstats, 0, sizeof(double) * I8_XSTATS));
  float ms_slice = time_ms([&] {
    cudaMemsetAsync(xstats, 0, sizeof(double) * I8_XSTATS);
    i8::slice_x_kernel<I8_TILE><<<(unsigned)std::min<int64_t>(n_tiles, sms * 8), 8 * I8_TILE>>>(x, ld, T, n, n_tiles, xblob, xstats);
  }, 1);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<double> hs(I8_XSTATS);
  CK(cudaMemcpy(hs.data(), xstats, sizeof(double) * I8_XSTATS, cudaMemcpyDeviceToHost));
  double min_ms = 1e300; for (int k = 0; k < n; ++k) min_ms = fmin(min_ms, hs[2 + k] / T);
  long long bits; memcpy(&bits, &hs[1], 8); double mx2; memcpy(&mx2, &bits, 8);
  printf("{\"kernel\": \"slice_x\", \"ms\": %.3f, \"mean_bound\": %.3f, \"min_row_rms\": %.4f, \"max_norm\": %.3f}\n", ms_slice, hs[0] / T, sqrt(min_ms), sqrt(mx2));

  if (only == 1) {
    run_loss<0>(xblob, w, y, ld, T, n, partial, sms, 1, d_trace);
    i8::row_exponent_kernel<<<16, 256>>>(w, n, xstats, rowexp);
    CK(cudaDeviceSynchronize());
    run_grad<0, 0>(y, ld, T, n, rowexp, partial, sms, 1, d_trace);
    printf("{\"done\": 1}\n");
    return 0;
  }
  if (only == 2) {
    // ---- full-size checks: the stored Y' of every sample against a naive product, run-to-run bit identity of both kernels, and the
    // gradient of the whole range (accumulator flushes every 512 tiles) against the sum of the gradients of eight sub-ranges (no flush)
    CK(cudaMalloc(&ynaive, sizeof(double) * n * ld));
    unsigned long long* d_out; CK(cudaMalloc(&d_out, 64));
    naive_y_kernel<<<(unsigned)(((int64_t)n * 8192 + 255) / 256), 256>>>(w, x, n, ld, 8192, ynaive);  // warm-up of the context
    CK(cudaDeviceSynchronize());
    unsigned long long first_sum = 0;
    for (int rep = 0; rep < reps; ++rep) {
      CK(cudaMemset(y, 0xff, sizeof(double) * n * ld));
      run_loss<0>(xblob, w, y, ld, T, n, partial, sms, 1, d_trace);
      CK(cudaDeviceSynchronize());
      CK(cudaMemset(d_out, 0, 64));
      checksum_kernel<<<sms * 8, 256>>>(y, ld, n, T, d_out);
      unsigned long long h[1]; CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
      if (rep == 0) first_sum = h[0];
      printf("{\"check\": \"loss_i8 Y' checksum\", \"rep\": %d, \"sum\": \"%016llx\", \"same_as_first\": %s}\n", rep, h[0], h[0] == first_sum ? "true" : "false");
    }
    {  // naive product of every sample, in slabs of 2^20 samples through the (n x tc) kernel
      const int64_t slab = 1 << 20;
      double* ys; CK(cudaMalloc(&ys, sizeof(double) * n * slab));
      unsigned long long tot_bad = 0; double worst = 0; long long first_bad = -1;
      for (int64_t c0 = 0; c0 < T; c0 += slab) {
        const int64_t tc = std::min<int64_t>(slab, T - c0);
        naive_y_kernel<<<(unsigned)(((int64_t)n * tc + 255) / 256), 256>>>(w, x + c0, n, ld, tc, ys);
        unsigned long long init[3] = {0, 0, ~0ull}; CK(cudaMemcpy(d_out, init, 24, cudaMemcpyHostToDevice));
        compare_kernel<<<sms * 8, 256>>>(ys, tc, y + c0, ld, n, tc, 1e-9, d_out);
        unsigned long long h[3]; CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
        if (h[0]) {
          double d; memcpy(&d, &h[1], 8);
          tot_bad += h[0]; worst = fmax(worst, d);
          if (first_bad < 0) { first_bad = (long long)(c0 + (long long)(h[2] % (unsigned long long)tc)); printf("{\"bad\": \"first in slab\", \"row\": %lld, \"sample\": %lld}\n", (long long)(h[2] / (unsigned long long)tc), first_bad); }
        }
      }
      printf("{\"check\": \"loss_i8 Y' vs naive f64, all samples\", \"T\": %lld, \"elements_off_by_more_than_1e-9\": %llu, \"worst\": %.3e}\n", (long long)T, tot_bad, worst);
      CK(cudaFree(ys));
    }
    i8::row_exponent_kernel<<<16, 256>>>(w, n, xstats, rowexp);
    CK(cudaDeviceSynchronize());
    const int n_tg = (int)std::min<int64_t>(sms / 2, (T + 31) / 32);
    const size_t psz = (size_t)2 * n_tg * (64 * 128 + 3 * 64);
    std::vector<double> first, g0, sd0;
    for (int rep = 0; rep < reps; ++rep) {
      CK(cudaMemset(partial, 0, sizeof(double) * psz));
      run_grad<0, 1>(y, ld, T, n, rowexp, partial, sms, 1, d_trace);
      CK(cudaDeviceSynchronize());
      std::vector<double> part(psz);
      CK(cudaMemcpy(part.data(), partial, psz * 8, cudaMemcpyDeviceToHost));
      size_t diff = 0;
      if (rep == 0) { first = part; reduce_grad(part, n_tg, n, g0, sd0); }
      else {
        const size_t per = 64 * 128 + 3 * 64;
        int imin = 1 << 30, imax = -1, jmin = 1 << 30, jmax = -1, smin = 1 << 30, smax = -1, cta_min = 1 << 30, cta_max = -1;
        double worst_g = 0, worst_s = 0;
        for (size_t i = 0; i < psz; ++i)
          if (memcmp(&part[i], &first[i], 8) != 0) {
            ++diff;
            const int cta = (int)(i / per); const size_t o = i % per;
            cta_min = std::min(cta_min, cta); cta_max = std::max(cta_max, cta);
            if (o < 64 * 128) {
              const int il = (int)(o / 128), j = (int)(o % 128);
              imin = std::min(imin, il); imax = std::max(imax, il); jmin = std::min(jmin, j); jmax = std::max(jmax, j);
              worst_g = fmax(worst_g, fabs(part[i] - first[i]));
            } else { const int k = (int)(o - 64 * 128); smin = std::min(smin, k); smax = std::max(smax, k); worst_s = fmax(worst_s, fabs(part[i] - first[i])); }
          }
        if (diff) printf("{\"diff\": \"where\", \"rep\": %d, \"cta\": [%d, %d], \"i_local\": [%d, %d], \"j\": [%d, %d], \"tail_index\": [%d, %d], \"worst_g\": %.3e, \"worst_tail\": %.3e}\n",
                         rep, cta_min, cta_max, imin, imax, jmin, jmax, smin, smax, worst_g, worst_s);
      }
      printf("{\"check\": \"grad_i8 partials, run-to-run\", \"rep\": %d, \"values_differing_from_first\": %zu}\n", rep, diff);
    }
    {
      const int chunks = 8;
      const int64_t cl = ((T / chunks + 31) / 32) * 32;
      std::vector<double> gs((size_t)n * n, 0.0), sds(n, 0.0);
      for (int64_t c0 = 0; c0 < T; c0 += cl) {
        const int64_t tc = std::min<int64_t>(cl, T - c0);
        const int ntg = (int)std::min<int64_t>(sms / 2, (tc + 31) / 32);
        CK(cudaMemset(partial, 0, sizeof(double) * psz));
        run_grad<0, 1>(y + c0, ld, tc, n, rowexp, partial, sms, 1, d_trace);
        CK(cudaDeviceSynchronize());
        std::vector<double> part((size_t)2 * ntg * (64 * 128 + 3 * 64)), g, sd;
        CK(cudaMemcpy(part.data(), partial, part.size() * 8, cudaMemcpyDeviceToHost));
        reduce_grad(part, ntg, n, g, sd);
        for (size_t i = 0; i < g.size(); ++i) gs[i] += g[i];
        for (int i = 0; i < n; ++i) sds[i] += sd[i];
      }
      double num = 0, den = 0, nsd = 0, dsd = 0;
      for (size_t i = 0; i < gs.size(); ++i) { num = fmax(num, fabs(g0[i] - gs[i])); den = fmax(den, fabs(gs[i])); }
      for (int i = 0; i < n; ++i) { nsd = fmax(nsd, fabs(sd0[i] - sds[i])); dsd = fmax(dsd, fabs(sds[i])); }
      printf("{\"check\": \"grad_i8 whole range vs sum of 8 sub-ranges\", \"T\": %lld, \"gr_rel\": %.3e, \"sd_rel\": %.3e}\n", (long long)T, num / den, nsd / dsd);
    }
    printf("{\"done\": 2}\n");
    return 0;
  }
  // ---- LOSS kernel: timing, ablations, trace
  run_loss<0>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);
  // correctness of the stored Y' on the first 8192 samples
  {
    const int64_t tc = std::min<int64_t>(T, 8192);
    CK(cudaMalloc(&ynaive, sizeof(double) * n * tc));
    naive_y_kernel<<<(unsigned)((n * tc + 255) / 256), 256>>>(w, x, n, ld, tc, ynaive);
    std::vector<double> a((size_t)n * tc), b((size_t)n * tc);
    CK(cudaMemcpy(a.data(), ynaive, sizeof(double) * n * tc, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy2D(b.data(), sizeof(double) * tc, y, sizeof(double) * ld, sizeof(double) * tc, n, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { num = fmax(num, fabs(a[i] - b[i])); den = fmax(den, fabs(a[i])); }
    printf("{\"check\": \"loss_i8 Y' vs naive f64\", \"max_abs_err\": %.3e, \"max_abs_y\": %.3f, \"rel\": %.3e}\n", num, den, num / den);
  }
  run_loss<1>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);
  run_loss<2>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);
  run_loss<3>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);
  run_loss<4>(xblob, w, y, ld, T, n, partial, sms, 1, d_trace);
  run_loss<8>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);   // no MMAs: what the epilogue alone costs
  run_loss<9>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);   // no MMAs, no density
  run_loss<11>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);  // no MMAs, no density, no store: tile hand-over only
  run_loss<0>(xblob, w, y, ld, T, n, partial, sms, reps, d_trace);  // leaves the full Y' for the gradient kernel

  // ---- gradient kernel: correctness of both layouts on a short prefix, then timing
  i8::row_exponent_kernel<<<16, 256>>>(w, n, xstats, rowexp);
  CK(cudaDeviceSynchronize());
  {
    const int64_t tc = std::min<int64_t>(T, 40000 + 7);
    CK(cudaMalloc(&gnaive, sizeof(double) * n * n)); CK(cudaMalloc(&sdnaive, sizeof(double) * n));
    naive_g_kernel<<<n, n>>>(y, n, ld, tc, gnaive, sdnaive);
    std::vector<double> gref((size_t)n * n), sdref(n);
    CK(cudaMemcpy(gref.data(), gnaive, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sdref.data(), sdnaive, sizeof(double) * n, cudaMemcpyDeviceToHost));
    for (int layout = 0; layout < 2; ++layout) {
      CK(cudaMemset(partial, 0, sizeof(double) * sms * (64 * 128 + 3 * 64)));
      if (layout == 0) run_grad<0, 0>(y, ld, tc, n, rowexp, partial, sms, 1, d_trace); else run_grad<0, 1>(y, ld, tc, n, rowexp, partial, sms, 1, d_trace);
      CK(cudaDeviceSynchronize());
      const int n_tg = (int)std::min<int64_t>(sms / 2, (tc + 31) / 32);
      std::vector<double> part((size_t)2 * n_tg * (64 * 128 + 3 * 64)), g, sd;
      CK(cudaMemcpy(part.data(), partial, part.size() * 8, cudaMemcpyDeviceToHost));
      reduce_grad(part, n_tg, n, g, sd);
      double num = 0, den = 0, nsd = 0, dsd = 0;
      for (size_t i = 0; i < g.size(); ++i) { num = fmax(num, fabs(g[i] - gref[i])); den = fmax(den, fabs(gref[i])); }
      for (int i = 0; i < n; ++i) { nsd = fmax(nsd, fabs(sd[i] - sdref[i])); dsd = fmax(dsd, fabs(sdref[i])); }
      printf("{\"check\": \"grad_i8 Gr vs naive f64\", \"layout\": %d, \"T\": %lld, \"rel\": %.3e, \"sd_rel\": %.3e, \"g00\": %.6f, \"ref00\": %.6f}\n", layout,
             (long long)tc, num / den, nsd / dsd, g[0], gref[0]);
    }
  }
  run_grad<0, 0>(y, ld, T, n, rowexp, partial, sms, reps, d_trace);
  run_grad<0, 1>(y, ld, T, n, rowexp, partial, sms, reps, d_trace);
  run_grad<1, 0>(y, ld, T, n, rowexp, partial, sms, reps, d_trace);
  run_grad<4, 0>(y, ld, T, n, rowexp, partial, sms, 1, d_trace);
  printf("{\"done\": 1}\n");
  return 0;
}
