// DMMA.8x8x4 / DFMA dependent-chain latency on B200: throughput vs number of independent accumulator chains
// per warp and warps per SM.  Tells how many independent chains step 1 of the pass kernel needs per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_latency profiles/microbench/dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
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
template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
  double* out; CK(cudaMalloc(&out, 64));
  const int iters = 20000;
#define RUN_DMMA(NACC, WARPS)                                                                                          \
  {                                                                                                                    \
    float ms = time_ms([&] { k_dmma<NACC><<<sms, WARPS * 32>>>(out, iters, 1.0000001, 1e-9); });                        \
    double cyc = ms * 1e-3 * ghz * 1e9; /* cycles */                                                                   \
    double per_round = cyc / iters;     /* cycles per round of NACC dependent-chain steps */                            \
    double tf = 2.0 * 256 * NACC * (double)iters * WARPS * sms / ms * 1e-9;                                              \
    printf("{\"test\": \"dmma_chain\", \"chains_per_warp\": %d, \"warps_per_sm\": %d, \"chains_per_smsp\": %.1f, \"cycles_per_round\": %.1f, \"tflops\": %.2f}\n", NACC, WARPS, NACC * WARPS / 4.0, per_round, tf); \
  }
  RUN_DMMA(1, 4) RUN_DMMA(2, 4) RUN_DMMA(4, 4) RUN_DMMA(8, 4) RUN_DMMA(16, 4)
  RUN_DMMA(1, 8) RUN_DMMA(2, 8) RUN_DMMA(4, 8) RUN_DMMA(8, 8) RUN_DMMA(16, 8)
  RUN_DMMA(1, 16) RUN_DMMA(2, 16) RUN_DMMA(4, 16) RUN_DMMA(8, 16)
  RUN_DMMA(1, 32) RUN_DMMA(2, 32) RUN_DMMA(4, 32)
#define RUN_DFMA(NCH, WARPS)                                                                                           \
  {                                                                                                                    \
    float ms = time_ms([&] { k_dfma<NCH><<<sms, WARPS * 32>>>(out, iters, 1.0000001, 1e-9); });                         \
    double cyc = ms * 1e-3 * ghz * 1e9;                                                                                \
    double tf = 2.0 * 32 * NCH * (double)iters * WARPS * sms / ms * 1e-9;                                               \
    printf("{\"test\": \"dfma_chain\", \"chains_per_thread\": %d, \"warps_per_sm\": %d, \"cycles_per_round\": %.1f, \"tflops\": %.2f}\n", NCH, WARPS, cyc / iters, tf); \
  }
  RUN_DFMA(1, 4) RUN_DFMA(2, 4) RUN_DFMA(4, 4) RUN_DFMA(8, 4) RUN_DFMA(1, 8) RUN_DFMA(2, 8) RUN_DFMA(4, 8) RUN_DFMA(8, 8) RUN_DFMA(4, 16) RUN_DFMA(8, 16)
  return 0;
}
