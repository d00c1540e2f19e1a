// profiles/lab/pipe_probe.cu -- does the INT8 tcgen05 MMA compete with the FP64 / FP32 / integer pipes of the SM?
// One CTA per SM.  Warp 0 (one lane) issues `n_mma` UTCIMMA (128 x N x 32, kind::i8, operands in shared memory or A in TMEM) into
// rotating TMEM accumulators and waits for their completion; warps 1..W run `iters` rounds of 8 independent chains of one ALU
// instruction kind (DFMA, FFMA, IMAD, or the mixed digit-conversion sequence).  Each role measures its own duration with clock64().
// Runs: MMA alone, ALU alone, both together.  Output: one JSON object per line.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../picard-ica_b200/csrc/i8_common.cuh"

using namespace picard;
using namespace picard::i8;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"cuda_error\": \"%s\", \"at\": \"%s\"}\n", cudaGetErrorString(e_), #x); exit(1); } } while (0)

template <int KIND>  // 0 DFMA, 1 FFMA, 2 IMAD, 3 DADD
__device__ __forceinline__ void alu_rounds(int iters, double* sink, int tid) {
  if (KIND == 0 || KIND == 3) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 1.0 + 1e-9 * (tid + i);
    const double a = 1.0000001, b = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = (KIND == 0) ? fma(c[i], a, b) : (c[i] + b);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
  } else if (KIND == 1) {
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 1.0f + 1e-3f * (tid + i);
    const float a = 1.0001f, b = 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123.456f) sink[0] = s;
  } else {
    int c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = tid + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = c[i] * 1664525 + 1013904223;
    }
    int s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123456) sink[0] = s;
  }
}

template <int KIND>
__global__ void __launch_bounds__(1024, 1) probe_kernel(int n_mma, int mma_n, int a_from_tmem, int iters, int alu_warps, long long* out, double* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) tmem_alloc512(&tmem_slot);
  ptx::fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (lane == 0 && n_mma > 0) {
      const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 16384));
      const uint32_t idesc = make_idesc(mma_n);
      const int nacc = a_from_tmem ? (448 / mma_n) : (512 / mma_n);  // rotating accumulators; A (32 columns) at column 448 when in TMEM
      // issue loop without integer division: 4 K-steps x rotating accumulators, everything but the loop counter precomputed
      uint32_t dcol[8];
      for (int a = 0; a < 8; ++a) dcol[a] = tmem + (uint32_t)((a % nacc) * mma_n);
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (a_from_tmem) umma_i8_ts(dcol[u], tmem + 448 + 8 * (u & 3), db + (uint64_t)(((u & 3) * 32) >> 4), idesc, 1u);
          else umma_i8_ss(dcol[u], da + (uint64_t)(((u & 3) * 32) >> 4), db + (uint64_t)(((u & 3) * 32) >> 4), idesc, 1u);
        }
      }
      umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      t1 = clock64();
      out[blockIdx.x * 4 + 0] = t1 - t0;
    }
  } else if (warp <= alu_warps && iters > 0) {
    t0 = clock64();
    alu_rounds<KIND>(iters, sink, tid);
    t1 = clock64();
    if (lane == 0 && warp == 1) out[blockIdx.x * 4 + 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tmem);
}

template <int KIND>
static void run(const char* name, int sms, int n_mma, int mma_n, int a_tmem, int iters, int alu_warps, long long* d_out, double* d_sink) {
  auto kern = probe_kernel<KIND>;
  const size_t smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(long long) * 4 * sms));
  for (int rep = 0; rep < 2; ++rep) kern<<<sms, 32 * (alu_warps + 1), smem>>>(n_mma, mma_n, a_tmem, iters, alu_warps, d_out, d_sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(4 * sms);
  CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
  double mma_cyc = 0, alu_cyc = 0;
  for (int b = 0; b < sms; ++b) { mma_cyc += (double)h[b * 4]; alu_cyc += (double)h[b * 4 + 1]; }
  mma_cyc /= sms; alu_cyc /= sms;
  // ALU: per SMSP, warps on it = alu_warps / 4 (warps 1..W round-robin over the 4 sub-partitions)
  const double alu_instr_per_warp = 8.0 * iters;
  printf("{\"alu\": \"%s\", \"mma_n\": %d, \"a_from_tmem\": %d, \"n_mma\": %d, \"alu_warps\": %d, \"iters\": %d, \"cycles_per_mma\": %.1f, "
         "\"alu_cycles_total\": %.0f, \"cycles_per_alu_warp_instr_per_smsp\": %.3f}\n",
         name, mma_n, a_tmem, n_mma, alu_warps, iters, n_mma ? mma_cyc / n_mma : 0.0, alu_cyc,
         iters ? alu_cyc / (alu_instr_per_warp * (alu_warps / 4.0)) : 0.0);
}

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d_out; double* d_sink;
  CK(cudaMalloc(&d_out, sizeof(long long) * 4 * sms)); CK(cudaMalloc(&d_sink, 64));
  const int NM = 4000;
  // MMA alone: rate vs N and operand source
  for (int n : {32, 64, 128, 192, 256}) for (int at = 0; at < 2; ++at) run<0>("none", sms, NM, n, at, 0, 16, d_out, d_sink);
  // ALU alone (16 warps = 4 per sub-partition), sized to last about as long as 4000 MMAs of N = 64 (~220k cycles)
  run<0>("dfma", sms, 0, 64, 0, 3400, 16, d_out, d_sink);
  run<3>("dadd", sms, 0, 64, 0, 3400, 16, d_out, d_sink);
  run<1>("ffma", sms, 0, 64, 0, 6800, 16, d_out, d_sink);
  run<2>("imad", sms, 0, 64, 0, 6800, 16, d_out, d_sink);
  // both together
  for (int n : {32, 64, 128, 256}) {
    run<0>("dfma", sms, NM, n, 1, 3400, 16, d_out, d_sink);
    run<1>("ffma", sms, NM, n, 1, 6800, 16, d_out, d_sink);
    run<2>("imad", sms, NM, n, 1, 6800, 16, d_out, d_sink);
  }
  run<0>("dfma", sms, NM, 64, 0, 3400, 16, d_out, d_sink);
  run<0>("dfma", sms, NM, 256, 0, 3400, 16, d_out, d_sink);
  run<0>("dfma", sms, NM, 256, 1, 13600, 16, d_out, d_sink);  // FP64 stream as long as the MMA stream
  run<0>("dfma", sms, NM, 64, 1, 850, 16, d_out, d_sink);   // light FP64 load (25 % of the pipe)
  run<0>("dfma", sms, NM, 64, 1, 3400, 4, d_out, d_sink);   // one warp per sub-partition
  printf("{\"done\": 1}\n");
  return 0;
}
