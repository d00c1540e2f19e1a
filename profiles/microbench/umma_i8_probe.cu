// Probe for the INT8 tensor path (tcgen05.mma kind::i8, accumulators in TMEM): D[128 x 64] (s32) = A[128 x 128] (s8, K-major)
// * B[64 x 128]^T (s8, K-major), operands in shared memory in the SWIZZLE_128B K-major layout, four K = 32 instructions.
// Self-checking against a CPU product; prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B, rows of 128 bytes: 8-row groups are 1024 bytes apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

constexpr int M = 128, N = 64, K = 128;
// instruction descriptor: c = s32 (2 << 4), a = b = signed int8 (1 << 7, 1 << 10), K-major both, n_dim = N / 8 at [17,23), m_dim = M / 16 at [24,29)
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128) probe(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sa = smem;                 // 128 rows x 128 B
  unsigned char* sb = smem + M * 128;       // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < M * K / 16; e += blockDim.x) {   // 16-byte chunks
    const int r = e / 8, ch = e % 8;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + ch * 16);
  }
  for (int e = tid; e < N * K / 16; e += blockDim.x) {
    const int r = e / 8, ch = e % 8;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + ch * 16);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy operand writes before the async-proxy MMA reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  uint32_t parity = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      const uint64_t da = make_desc(smem_u32(sa)), db = make_desc(smem_u32(sb));
#pragma unroll
      for (int k = 0; k < K / 32; ++k) {
        const uint32_t acc = k > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
            "}\n" ::"r"(tmem),
            "l"(da + (uint64_t)((k * 32) >> 4)), "l"(db + (uint64_t)((k * 32) >> 4)), "r"(IDESC), "r"(acc), "r"(0u)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(&bar)),
        "r"(parity)
        : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // D row = TMEM lane (warp w owns lanes 32 w .. 32 w + 31), D column = TMEM column
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]), "=r"(v[c0 + 6]),
          "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]), "=r"(v[c0 + 12]), "=r"(v[c0 + 13]),
          "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
        : "r"(taddr + c0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; ++c) D[row * N + c] = (int32_t)v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// A operand from TENSOR MEMORY (tcgen05.mma ... [tmem_a]): row i of A in TMEM lane i, the 32 K-bytes of one K = 32 step in 8
// consecutive 32-bit columns (4 bytes per column, K ascending) -- the layout this probe verifies.  D = 64 columns at tmem + 0,
// A at tmem + 64 + 8 k.
__global__ void __launch_bounds__(128) probe_ts(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sb = smem;                 // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < N * K / 16; e += blockDim.x) {
    const int r = e / 8, ch = e % 8;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + ch * 16);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // every thread stores its row of A: 4 K-steps x 8 words
  const int row = warp * 32 + lane;
  const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + row * K);
#pragma unroll
  for (int k = 0; k < K / 32; ++k) {
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 64 + 8 * k;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(arow[8 * k + 0]),
                 "r"(arow[8 * k + 1]), "r"(arow[8 * k + 2]), "r"(arow[8 * k + 3]), "r"(arow[8 * k + 4]), "r"(arow[8 * k + 5]),
                 "r"(arow[8 * k + 6]), "r"(arow[8 * k + 7])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t db = make_desc(smem_u32(sb));
#pragma unroll
    for (int k = 0; k < K / 32; ++k) {
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
          "}\n" ::"r"(tmem),
          "r"(tmem + 64 + 8 * k), "l"(db + (uint64_t)((k * 32) >> 4)), "r"(IDESC), "r"(acc), "r"(0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(&bar)),
      "r"(0u)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]), "=r"(v[c0 + 6]),
          "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]), "=r"(v[c0 + 12]), "=r"(v[c0 + 13]),
          "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
        : "r"(taddr + c0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 64; ++c) D[row * N + c] = (int32_t)v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// Pipelined issue rate: `batch` MMAs of 128 x NN x 32 back to back (alternating between two accumulators), one commit per batch.
template <int NN, bool TS>
__global__ void __launch_bounds__(128) rate(int batches, int batch) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (M + 64) * 128 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t parity = 0;
  for (int b = 0; b < batches; ++b) {
    if (tid == 0) {
      const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + M * 128));
      for (int i = 0; i < batch; ++i) {
        if (!TS)
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
              "}\n" ::"r"(tmem + (uint32_t)((i & 1) * 64)),
              "l"(da + (uint64_t)(((i & 3) * 32) >> 4)), "l"(db + (uint64_t)(((i & 3) * 32) >> 4)), "r"(idesc), "r"((uint32_t)(i > 1)), "r"(0u)
              : "memory");
        else
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
              "}\n" ::"r"(tmem + (uint32_t)((i & 1) * 64)),
              "r"(tmem + 128 + (uint32_t)((i & 3) * 8)), "l"(db + (uint64_t)(((i & 3) * 32) >> 4)), "r"(idesc), "r"((uint32_t)(i > 1)), "r"(0u)
              : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(&bar)),
        "r"(parity)
        : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int NN, bool TS>
static void run_rate(int sms) {
  const int smem = (M + 64) * 128 + 1024, batches = 2000, batch = 112;
  cudaFuncSetAttribute(rate<NN, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<NN, TS><<<sms, 128, smem>>>(10, batch);
  cudaEventRecord(e0); rate<NN, TS><<<sms, 128, smem>>>(batches, batch); cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double per_mma_ns = ms * 1e6 / ((double)batches * batch);
  const double tops = 2.0 * M * NN * 32 * (double)batches * batch * sms / (ms * 1e-3) / 1e12;
  printf("{\"test\": \"umma_i8_rate\", \"a_from_tmem\": %d, \"n\": %d, \"cuda\": \"%s\", \"ns_per_mma\": %.2f, \"cycles_per_mma_at_1965\": %.1f, \"tops\": %.1f}\n", (int)TS, NN,
         cudaGetErrorString(e), per_mma_ns, per_mma_ns * 1.965, tops);
}

int main() {
  std::vector<int8_t> a(M * K), b(N * K);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 16) % 255) - 127; };
  for (auto& x : a) x = (int8_t)rnd();
  for (auto& x : b) x = (int8_t)rnd();
  std::vector<int32_t> ref(M * N, 0), got(M * N, -1);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t acc = 0;
      for (int k = 0; k < K; ++k) acc += (int32_t)a[i * K + k] * (int32_t)b[j * K + k];
      ref[i * N + j] = acc;
    }
  int8_t *da, *db; int32_t* dd;
  cudaMalloc(&da, a.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dd, got.size() * 4);
  cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xFF, got.size() * 4);
  const int smem = (M + N) * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(da, db, dd, 1);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0; int first = -1;
  for (int i = 0; i < M * N; ++i) if (got[i] != ref[i]) { if (first < 0) first = i; ++bad; }
  // throughput: one CTA per SM, many repetitions of the 4-instruction group
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int reps = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<<<sms, 128, smem>>>(da, db, dd, 100);
  cudaEventRecord(e0); probe<<<sms, 128, smem>>>(da, db, dd, reps); cudaEventRecord(e1);
  cudaError_t e2 = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double tops = 2.0 * M * N * K * (double)reps * sms / (ms * 1e-3) / 1e12;
  printf("{\"test\": \"umma_i8_probe\", \"cuda\": \"%s / %s\", \"mismatches\": %ld, \"first\": %d, \"got0\": %d, \"ref0\": %d, \"ms\": %.3f, \"tops_serialised_groups\": %.1f}\n",
         cudaGetErrorString(e), cudaGetErrorString(e2), bad, first, got[0], ref[0], ms, tops);
  {
    cudaMemset(dd, 0xFF, got.size() * 4);
    const int smem_ts = N * 128 + 1024;
    cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ts);
    probe_ts<<<1, 128, smem_ts>>>(da, db, dd);
    cudaError_t ets = cudaDeviceSynchronize();
    std::vector<int32_t> got2(M * N, -1);
    cudaMemcpy(got2.data(), dd, got2.size() * 4, cudaMemcpyDeviceToHost);
    long bad2 = 0; int first2 = -1;
    for (int i = 0; i < M * N; ++i) if (got2[i] != ref[i]) { if (first2 < 0) first2 = i; ++bad2; }
    printf("{\"test\": \"umma_i8_a_from_tmem\", \"cuda\": \"%s\", \"mismatches\": %ld, \"first\": %d, \"got0\": %d, \"ref0\": %d, \"got1\": %d, \"ref1\": %d}\n",
           cudaGetErrorString(ets), bad2, first2, got2[0], ref[0], got2[N], ref[N]);
  }
  run_rate<32, false>(sms); run_rate<64, false>(sms); run_rate<16, true>(sms); run_rate<32, true>(sms); run_rate<64, true>(sms);
  return bad != 0;
}
