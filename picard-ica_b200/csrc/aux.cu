// Auxiliary HBM-bound kernels: row sums for centering (whitening.rs:24-35) and the counter-based synthetic
// source generator of the benchmark.
#include <cmath>

#include "engine.cuh"

namespace picard {
namespace aux {

namespace {
constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS) row_sum_stage1(const double* __restrict__ x, int64_t t_local, int64_t ldx, int nchunks,
                                                             double* __restrict__ work) {
  __shared__ double sh[RS_THREADS / 32];
  const int row = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (t_local + nchunks - 1) / nchunks;
  const int64_t a = chunk * per, b = (a + per < t_local) ? a + per : t_local;
  const double* p = x + (size_t)row * ldx;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int64_t i = a + threadIdx.x;
  for (; i + 3 * RS_THREADS < b; i += 4 * RS_THREADS) {
    s0 += p[i]; s1 += p[i + RS_THREADS]; s2 += p[i + 2 * RS_THREADS]; s3 += p[i + 3 * RS_THREADS];
  }
  for (; i < b; i += RS_THREADS) s0 += p[i];
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0;
    for (int w = 0; w < RS_THREADS / 32; ++w) r += sh[w];
    work[(size_t)row * nchunks + chunk] = r;
  }
}
__global__ void row_sum_stage2(const double* __restrict__ work, int nchunks, double* __restrict__ out) {
  __shared__ double sh[8];
  const int row = blockIdx.x;
  double s = 0;
  for (int c = threadIdx.x; c < nchunks; c += blockDim.x) s += work[(size_t)row * nchunks + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
    out[row] = r;
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// S[i, t]: u = ((mix64(key_i + (t+1) * GOLDEN) >> 11) + 0.5) * 2^-53 in (0,1), key_i = mix64(seed ^ mix64(i + 0x1234567));
// Laplace(b = 1/sqrt 2) by inverse CDF for i < n_laplace, else uniform on [-sqrt 3, sqrt 3].
__global__ void synth_kernel(double* __restrict__ out, int n, int64_t t_local, int64_t ld, int64_t t_offset, int n_laplace,
                             uint64_t seed) {
  const int row = blockIdx.y;
  const uint64_t key = mix64(seed ^ mix64((uint64_t)row + 0x1234567ull));
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < t_local; s += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t h = mix64(key + (uint64_t)(t_offset + s + 1) * 0x9E3779B97F4A7C15ull);
    const double u = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double v;
    if (row < n_laplace) {
      const double d = u - 0.5;
      const double b = 0.70710678118654752440;
      v = d < 0 ? b * log(1.0 + 2.0 * d) : -b * log(1.0 - 2.0 * d);
    } else {
      v = 1.7320508075688772935 * (2.0 * u - 1.0);
    }
    out[(size_t)row * ld + s] = v;
  }
}
}  // namespace

int row_sums(const double* d_x, int n, int64_t t_local, int64_t ldx, double* d_work, double* d_out, cudaStream_t st) {
  int nchunks = (int)((t_local + 8191) / 8192);
  if (nchunks > 1024) nchunks = 1024;
  if (nchunks < 1) nchunks = 1;
  row_sum_stage1<<<dim3(nchunks, n), RS_THREADS, 0, st>>>(d_x, t_local, ldx, nchunks, d_work);
  PICARD_CUDA(cudaGetLastError());
  row_sum_stage2<<<n, 256, 0, st>>>(d_work, nchunks, d_out);
  PICARD_CUDA(cudaGetLastError());
  return 2;
}

int synth_sources(double* d_out, int n, int64_t t_local, int64_t ld, int64_t t_offset, int n_laplace, uint64_t seed, cudaStream_t st) {
  int64_t bx = (t_local + 1023) / 1024;
  if (bx > 2048) bx = 2048;
  if (bx < 1) bx = 1;
  synth_kernel<<<dim3((unsigned)bx, n), 256, 0, st>>>(d_out, n, t_local, ld, t_offset, n_laplace, seed);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}


// FP64 tensor peak of THIS device at its CURRENT clocks: every warp runs eight independent DMMA.8x8x4 accumulator chains
// (the pipe saturates with two chains per scheduler; profiles/microbench/dmma_latency_r01.jsonl).  bench.py runs it right
// before the timed region and uses the result as the roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
namespace {
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* __restrict__ sink) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) ptx::dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}
}  // namespace

double fp64_peak_probe(int sm_count, double budget_ms, cudaStream_t st) {
  DevBuf<double> sink(1);
  cudaEvent_t e0, e1;
  PICARD_CUDA(cudaEventCreate(&e0)); PICARD_CUDA(cudaEventCreate(&e1));
  const int grid = sm_count * 4;
  int iters = 20000;
  double best = 0.0, spent = 0.0;
  for (int rep = 0; rep < 12 && spent < budget_ms; ++rep) {
    PICARD_CUDA(cudaEventRecord(e0, st));
    dmma_probe_kernel<<<grid, 256, 0, st>>>(iters, sink.p);
    PICARD_CUDA(cudaEventRecord(e1, st));
    PICARD_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    PICARD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    spent += ms;
    const double flops = (double)grid * 8.0 /* warps */ * 8.0 /* chains */ * (double)iters * 512.0;  // m8n8k4: 256 FMA per warp instruction
    if (rep > 0) best = std::fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}
}  // namespace aux
}  // namespace picard
