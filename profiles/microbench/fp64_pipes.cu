// FP64 pipe microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries no FP64 figure, and the design of the fused moments pass depends on
// three questions (SURVEY.md §2.4, §7 "Hard parts"):
//   1. what is the DFMA (vector FP64) peak,
//   2. what is the DMMA (mma.sync f64) peak, per shape,
//   3. do DMMA and DFMA issue to the same pipe (mixed kernels: does time add or overlap)?
// Also: cublasDgemm 8192^3 and a read-only HBM streaming figure.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo fp64_pipes.cu -lcublas -o fp64_pipes
// Output: one JSON line per measurement on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- DFMA only: NCH independent chains per thread
template <int NCH>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

// ---- DMMA m8n8k4 only: NACC independent accumulator tiles per warp
template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
  double c[NACC][4], a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
#pragma unroll
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
  double c[NACC][4], a[8], b[4];
  for (int j = 0; j < 8; ++j) a[j] = a0 + j;
  for (int j = 0; j < 4; ++j) b[j] = b0 + j;
#pragma unroll
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456) out[0] = s;
}

// ---- mixed in the same warp: NACC DMMA + NF DFMA per loop trip
template <int NACC, int NF>
__global__ void k_mixed(double* out, int iters, double a, double b) {
  double c[NACC][2], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

// ---- mixed across warps: even warps DMMA, odd warps DFMA
__global__ void k_split(double* out, int iters_mma, int iters_fma, double a, double b) {
  int w = threadIdx.x >> 5;
  double s = 0;
  if (w & 1) {
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters_fma; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
  } else {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
    for (int it = 0; it < iters_mma; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  }
  if (s == 123.456) out[0] = s;
}

// ---- transcendental cost: libdevice exp/log/div per element
__global__ void k_transc(double* out, int iters, double x0, int mode) {
  double x = x0 + threadIdx.x * 1e-4, acc = 0;
  for (int it = 0; it < iters; ++it) {
    double v;
    if (mode == 0) v = exp(-2.0 * fabs(x));
    else if (mode == 1) v = log(1.0 + x * x);
    else if (mode == 2) v = 1.0 / (1.0 + x * x);
    else v = tanh(x);
    acc += v; x += 1e-6;
  }
  if (acc == 123.456) out[0] = acc;
}

// ---- HBM streaming read
__global__ void k_stream(const double2* __restrict__ x, size_t n2, double* out) {
  double s = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n2; i += 4 * st) {
    double2 a = x[i], b = x[i + st], c = x[i + 2 * st], d = x[i + 3 * st];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += st) { double2 a = x[i]; s += a.x + a.y; }
  if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, 64));
  const int TPB = 256;
  // warps per SM sweep: blocks per SM x 8 warps
  for (int bps : {1, 2, 4, 8}) {
    int grid = sms * bps; int iters = 20000;
    float ms = time_ms([&] { k_dfma<8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 8 * iters * (double)TPB * grid;
    printf("{\"test\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
  }
  for (int bps : {1, 2, 4, 8}) {
    int grid = sms * bps; int iters = 20000;
    float ms = time_ms([&] { k_dmma884<8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 256 * 8 * iters * (double)(TPB / 32) * grid;
    printf("{\"test\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
  }
  for (int bps : {1, 2, 4}) {
    int grid = sms * bps; int iters = 10000;
    float ms = time_ms([&] { k_dmma1688<8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)(TPB / 32) * grid;
    printf("{\"test\": \"dmma_m16n8k8\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
    ms = time_ms([&] { k_dmma16816<8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 16 * 8 * 16 * 8 * iters * (double)(TPB / 32) * grid;
    printf("{\"test\": \"dmma_m16n8k16\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
  }
  // mixed, same warp: 8 DMMA (=8*256 FMA/warp = 64 FMA/thread-equiv) + NF DFMA
  {
    int grid = sms * 2, iters = 20000;
    float t_m = time_ms([&] { k_mixed<8, 0><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    float t_8 = time_ms([&] { k_mixed<8, 8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    float t_16 = time_ms([&] { k_mixed<8, 16><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    float t_32 = time_ms([&] { k_mixed<8, 32><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    float t_64 = time_ms([&] { k_mixed<8, 64><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    float f_8 = time_ms([&] { k_dfma<8><<<grid, TPB>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"test\": \"mixed_same_warp\", \"ms_dmma8_only\": %.4f, \"ms_dfma8_only\": %.4f, \"ms_dmma8_dfma8\": %.4f, "
           "\"ms_dmma8_dfma16\": %.4f, \"ms_dmma8_dfma32\": %.4f, \"ms_dmma8_dfma64\": %.4f}\n", t_m, f_8, t_8, t_16, t_32, t_64);
  }
  {
    int grid = sms * 2;
    // even warps: 20000*8 DMMA; odd warps: iters_fma*8 DFMA. Pick DFMA count so that each alone takes similar time.
    float t_mma = time_ms([&] { k_split<<<grid, TPB>>>(out, 20000, 0, 1.0000001, 1e-9); });
    float t_fma = time_ms([&] { k_split<<<grid, TPB>>>(out, 0, 160000, 1.0000001, 1e-9); });
    float t_both = time_ms([&] { k_split<<<grid, TPB>>>(out, 20000, 160000, 1.0000001, 1e-9); });
    printf("{\"test\": \"mixed_split_warps\", \"ms_dmma_half_warps\": %.4f, \"ms_dfma_half_warps\": %.4f, \"ms_both\": %.4f}\n",
           t_mma, t_fma, t_both);
  }
  for (int mode = 0; mode < 4; ++mode) {
    int grid = sms * 4, iters = 4000;
    float ms = time_ms([&] { k_transc<<<grid, TPB>>>(out, iters, 0.3, mode); });
    double el = (double)iters * TPB * grid;
    const char* nm[] = {"exp", "log", "div", "tanh"};
    printf("{\"test\": \"libdevice_%s\", \"ms\": %.4f, \"gelem_per_s\": %.2f}\n", nm[mode], ms, el / ms * 1e-6);
  }
  {
    size_t n = (size_t)1 << 29;  // 4 GiB of doubles
    double* x; CK(cudaMalloc(&x, n * 8)); CK(cudaMemset(x, 0, n * 8));
    float ms = time_ms([&] { k_stream<<<sms * 8, 512>>>((const double2*)x, n / 2, out); });
    printf("{\"test\": \"hbm_stream_read\", \"bytes\": %zu, \"ms\": %.4f, \"gbs\": %.1f}\n", n * 8, ms, n * 8.0 / ms * 1e-6);
    CK(cudaFree(x));
  }
  {
    int n = 8192; double *A, *B, *C;
    CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
    CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
    cublasHandle_t h; cublasCreate(&h); double al = 1, be = 0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
    printf("{\"test\": \"cublas_dgemm_8192\", \"ms\": %.3f, \"tflops\": %.3f}\n", ms, 2.0 * n * n * (double)n / ms * 1e-9);
    // the shape of our contraction: (128 x K) x (K x 128), K = 1e6
    int m = 128, k = 1000000;
    double *P, *Q, *R; CK(cudaMalloc(&P, (size_t)m * k * 8)); CK(cudaMalloc(&Q, (size_t)m * k * 8)); CK(cudaMalloc(&R, (size_t)m * m * 8));
    CK(cudaMemset(P, 0, (size_t)m * k * 8)); CK(cudaMemset(Q, 0, (size_t)m * k * 8));
    ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, m, m, k, &al, P, k, Q, k, &be, R, m); }, 3);
    printf("{\"test\": \"cublas_dgemm_128x128xK1e6_TN\", \"ms\": %.3f, \"tflops\": %.3f}\n", ms, 2.0 * m * m * (double)k / ms * 1e-9);
    ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, k, m, m, &al, P, k, R, m, &be, Q, k); }, 3);
    printf("{\"test\": \"cublas_dgemm_K1e6x128x128_NN\", \"ms\": %.3f, \"tflops\": %.3f}\n", ms, 2.0 * m * m * (double)k / ms * 1e-9);
  }
  return 0;
}
