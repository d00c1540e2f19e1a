// =====================================================================================================
// i8_loss.cu -- the LOSS pass (Y' = W' x1, log-likelihood / y^2 row sums, Y' kept in HBM: core.rs:124-127 with compute_loss
// core.rs:39-85) on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) instead of the FP64 DMMA path.
// Used for 64 < N <= 128 when the data is whitened (every component of a sample is O(1), so the per-sample scaling loses
// nothing); PICARD_I8=0 selects the FP64 path (rb_loss_kernel, rowblock.cuh) everywhere, PICARD_I8=1 forces this one.
//
// Error-free splitting (Ozaki-type).  Every row of W' and every sample (column) of x1 is scaled by its own power of two into
// (-1, 1) and cut into S = 7 signed 7-bit slices (q_p = trunc(128 r_p), r_{p+1} = 128 r_p - q_p: exact in f64), so
//   w'_ik x_kt = 2^(ew_i + ex_t) sum_{p,q} a_p b_q 128^-(p+q+2).
// The slice products with p + q = d <= S - 1 are INT8 GEMMs whose sums are exact in the s32 accumulator of their level d
// (|sum| <= (d + 1) 128 127^2 < 2^24); the dropped products are below 2^-49 of max|w'_i.| max|x_.t| -- measured 1.7e-13 of
// max|y| against exact rational arithmetic (tools/ozaki_numerics.py; f64 dgemm: 3.9e-16), three orders inside the parity bar
// of 1e-10.  The 7 levels are combined exactly in 64-bit integers (two halves), converted once and rounded once.
//
// x1 is fixed for a whole fit, so it is sliced ONCE (i8_slice_x) into tiles of 32 samples, each tile already in the shared-
// memory image the tensor core reads (K-major rows of 128 bytes, SWIZZLE_128B pattern) followed by the 32 column scales: a tile
// is one contiguous 28.25 KB bulk copy (cp.async.bulk, no tensor map).  W' (128 x 128) is sliced per try by a one-CTA kernel.
//
// loss_i8_kernel, one CTA per SM, 18 warps:
//   warp 16 (one lane) : bulk copies of the x1 tiles into a 2-stage ring (mbarrier transaction bytes)
//   warp 17 (one lane) : per tile 28 products x 4 K-steps = 112 tcgen05.mma (128 x 32 x 32, kind::i8).  The A operand (the W'
//                        slices, 224 columns) lives in TENSOR MEMORY next to the 7 level accumulators (7 x 32 columns), so the
//                        MMAs only read the 1 KB B operand from shared memory; tcgen05.commit releases the ring stage and
//                        publishes the accumulators
//   warps 0-15         : epilogue, thread = (row = TMEM lane, 8 samples): tcgen05.ld of the 7 levels, accumulators returned at
//                        once, exact 64-bit combination, scaling, log-likelihood (density.cuh, 80 KB tables), row sums per
//                        thread; Y' through shared memory and ONE TMA store per 16 samples (direct stores of a thread's 64
//                        bytes touch 32 lines per instruction: 9.3 -> 8.2 ms); warps 0-3 first store W' into tensor memory
// Measured (B200, N = 128, T = 1e7, tanh, profiles/microbench/umma_i8_probe.cu and pass_bench.py with PICARD_I8=1):
//   * a 128 x N x 32 kind::i8 MMA costs 55 cycles whatever N <= 64 (51 with A from tensor memory), so with N = 32 (two
//     accumulator sets or W' must fit the 512 TMEM columns) the pass is bound by MMA issue at 112 x 28 ns per 32 samples = 6.6 ms;
//   * the pass takes 8.2 ms (rb_loss_kernel on the FP64 DMMA path: 10.85 ms); without density and store 6.9 ms.
// =====================================================================================================
#include "i8_loss.cuh"

#include "rowblock_inst.cuh"

namespace picard {

namespace i8 {

constexpr int S = I8_SLICES;            // slices per operand
constexpr int KP = 128;                 // padded contraction length = bytes per operand row
constexpr int NT = I8_TILE;             // samples per tile
constexpr int SLICE_B_BYTES = NT * KP;  // 4096
constexpr int TILE_BYTES = I8_TILE_BYTES;
constexpr int STAGE_BYTES = ((TILE_BYTES + 1023) / 1024) * 1024;
constexpr int SLICE_A_BYTES = 128 * KP;  // 16384
constexpr int NSTAGE = 2;
constexpr int ACC_COLS = S * NT;         // 224 TMEM columns: the 7 level accumulators of one tile
#ifdef I8_TWO_GROUPS   // A/B build: the level accumulators handed over in two groups (levels 0-4: 60 MMAs, 5-6: 52 MMAs)
constexpr int NGROUP = 2;
#else
constexpr int NGROUP = 1;
#endif
__host__ __device__ constexpr int group_begin(int g) { return NGROUP == 1 ? (g == 0 ? 0 : S) : (g == 0 ? 0 : (g == 1 ? 5 : S)); }
constexpr int TMEM_A = ACC_COLS;         // W' slices live in tensor memory too: slice p, K-step k at column TMEM_A + 8 (4 p + k)
constexpr int NEW = 16;                  // epilogue warps: TMEM lane quarter (warp & 3) x column quarter of the tile (warp >> 2)
constexpr int CPT = NT / (NEW / 4);      // samples per epilogue thread and tile (8)
constexpr int NTHREADS = 32 * (NEW + 2);
constexpr bool BIG = true;               // density tables (80 KB): shared memory only holds the x1 ring besides
constexpr size_t SMEM_Y = (size_t)2 * 128 * NT * 8;            // Y' tile, two buffers of two [128 rows][16 samples] SWIZZLE_128B boxes
constexpr size_t SMEM_B = (size_t)NSTAGE * STAGE_BYTES;        // 59392
constexpr size_t SMEM_BYTES = SMEM_Y + SMEM_B + (size_t)dmath::Tab<BIG>::DOUBLES * 8 + 8192 /* sums */ + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows of 128 bytes; 8-row groups 1024 bytes apart (SBO); LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = NT, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// D[tmem_d] (+)= A[tmem_a] * B[smem descriptor]: A (128 rows x 32 K-bytes) in tensor memory, row = lane, 8 columns of 4 bytes
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 8 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// v 2^k on the ALU (v = 0 or a normal double far from the exponent limits): every FP64-pipe instruction of the epilogue competes
// with the tensor core for the pipe (ncu: 57 % of the epilogue's warp samples wait on the math pipe while it is 15 % active)
__device__ __forceinline__ double scale_pow2(double v, int k) {
  const int h = __double2hiint(v), l = __double2loint(v);
  return (((h << 1) | l) != 0) ? __hiloint2double(h + (k << 20), l) : v;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// slicing of x1: one CTA per tile of 32 samples, thread = (sample r, 16-byte chunk ch of its 128 slice bytes)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) slice_x_kernel(const double* __restrict__ x, int64_t ldx, int64_t t_local, int n,
                                                      uint8_t* __restrict__ blob) {
  const int tid = threadIdx.x, r = tid >> 3, ch = tid & 7;
  const int64_t tile = blockIdx.x, t = tile * NT + r;
  double v[16];
  double m = 0.0;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int k = 16 * ch + kk;
    v[kk] = (k < n && t < t_local) ? x[(size_t)k * ldx + t] : 0.0;
    m = fmax(m, fabs(v[kk]));
  }
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4));
  // 2^e > m >= 2^(e-1): x / 2^e lies in (-1, 1)
  int e = 0;
  if (m > 0.0) e = ilogb(m) + 1;
  const double inv = scalbn(1.0, -e);
  uint8_t* tb = blob + (size_t)tile * TILE_BYTES;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) v[kk] *= inv;
#pragma unroll
  for (int p = 0; p < S; ++p) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      v[kk] *= 128.0;
      const double q = trunc(v[kk]);
      v[kk] -= q;
      w[kk >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)q) << (8 * (kk & 3));
    }
    *reinterpret_cast<uint4*>(tb + (size_t)p * SLICE_B_BYTES + r * KP + ((ch ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (ch == 0) reinterpret_cast<double*>(tb + (size_t)S * SLICE_B_BYTES)[r] = scalbn(1.0, e);
}

// slicing of W' (n_out x n_in, zero-padded to 128 x 128): one CTA, thread = (row, chunk); plain row-major slices (they go to
// tensor memory row by row); the row scales already carry the 2^-28 of the level combination
__global__ void __launch_bounds__(1024) slice_w_kernel(const double* __restrict__ w, int ldw, int n_out, int n_in,
                                                       uint8_t* __restrict__ wblob) {
  const int tid = threadIdx.x, r = tid >> 3, ch = tid & 7;
  double v[16];
  double m = 0.0;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int k = 16 * ch + kk;
    v[kk] = (r < n_out && k < n_in) ? w[(size_t)r * ldw + k] : 0.0;
    m = fmax(m, fabs(v[kk]));
  }
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4));
  int e = 0;
  if (m > 0.0) e = ilogb(m) + 1;
  const double inv = scalbn(1.0, -e);
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) v[kk] *= inv;
#pragma unroll
  for (int p = 0; p < S; ++p) {
    uint32_t q4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      v[kk] *= 128.0;
      const double q = trunc(v[kk]);
      v[kk] -= q;
      q4[kk >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)q) << (8 * (kk & 3));
    }
    *reinterpret_cast<uint4*>(wblob + (size_t)p * SLICE_A_BYTES + r * KP + (ch << 4)) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
  }
  if (ch == 0) reinterpret_cast<double*>(wblob + (size_t)S * SLICE_A_BYTES)[r] = scalbn(1.0, e - 28);
}

// ---------------------------------------------------------------------------------------------------
// the pass
// ---------------------------------------------------------------------------------------------------
template <int DENS, bool WANT_SQ>
__global__ void __launch_bounds__(NTHREADS, 1)
loss_i8_kernel(const uint8_t* __restrict__ xblob, const uint8_t* __restrict__ wblob, const __grid_constant__ CUtensorMap tmap_out,
               const PassParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  double* ysm = reinterpret_cast<double*>(smem);
  unsigned char* sb = smem + SMEM_Y;
  double* tab = reinterpret_cast<double*>(smem + SMEM_Y + SMEM_B);
  double* sums = tab + dmath::Tab<BIG>::DOUBLES;  // [NEW / 4][2][128]: partial row sums of the column quarters
  uint64_t* bars = reinterpret_cast<uint64_t*>(sums + 1024);
  uint64_t* a_full = bars;                 // W' slices stored in tensor memory (4 warps)
  uint64_t* b_full = bars + 1;             // NSTAGE
  uint64_t* b_empty = b_full + NSTAGE;     // NSTAGE : MMA commit + the 8 epilogue warps (they read the column scales)
  // ONE accumulator set (there is no room for two next to the W' slices in tensor memory): the epilogue reads it back at once and
  // returns it before it combines the levels.  Handing the levels over in two groups (-DI8_TWO_GROUPS) was measured slower twice
  // (9.8 vs 9.3 ms with direct stores, 8.66 vs 8.39 ms with the TMA-store path): the pass is bound by its epilogue, not by the hand-over.
  uint64_t* acc_full = b_empty + NSTAGE;   // [NGROUP] accumulators (of a level group) of a tile complete (MMA commit)
  uint64_t* acc_empty = acc_full + 2;      // [NGROUP] accumulators read back (NEW epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool NEED_TAB = (DENS == DENS_TANH || DENS == DENS_EXP);
  if (NEED_TAB) load_density_tables<BIG>(tab, DENS == DENS_TANH, tid, NTHREADS);
  if (tid == 0) {
    ptx::mbar_init(a_full, 4);
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1 + NEW); }
    for (int g = 0; g < NGROUP; ++g) { ptx::mbar_init(&acc_full[g], 1); ptx::mbar_init(&acc_empty[g], NEW); }
    ptx::fence_barrier_init();
  }
  if (warp == NEW + 1) {  // the MMA warp owns the tensor memory: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t tile0 = blockIdx.x, tstride = gridDim.x;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;

  if (warp == NEW) {
    // =================================== producer ===================================
    if (lane == 0) {
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % NSTAGE);
        ptx::mbar_wait(&b_empty[st], (uint32_t)(((it / NSTAGE) & 1) ^ 1));  // first round: passes on the fresh barrier
        ptx::mbar_expect_tx(&b_full[st], (uint32_t)TILE_BYTES);
        bulk_load(sb + (size_t)st * STAGE_BYTES, xblob + (size_t)(tile0 + it * tstride) * TILE_BYTES, TILE_BYTES, &b_full[st]);
      }
    }
  } else if (warp == NEW + 1) {
    // =================================== MMA issue ===================================
    if (lane == 0) {
      ptx::mbar_wait(a_full, 0);
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % NSTAGE);
        ptx::mbar_wait(&b_full[st], (uint32_t)((it / NSTAGE) & 1));
        const uint64_t db0 = make_desc(smem_u32(sb + (size_t)st * STAGE_BYTES));
#pragma unroll
        for (int g = 0; g < NGROUP; ++g) {
          ptx::mbar_wait(&acc_empty[g], (uint32_t)((it & 1) ^ 1));  // the epilogue has read this group of tile it - 1 back (first tile: passes)
          tc_fence_after();
#pragma unroll
          for (int d = group_begin(g); d < group_begin(g + 1); ++d) {
            const uint32_t tacc = tmem + (uint32_t)(d * NT);
#pragma unroll
            for (int pa = 0; pa <= d; ++pa) {
              const int qb = d - pa;
#pragma unroll
              for (int k = 0; k < KP / 32; ++k)
                umma_i8(tacc, tmem + (uint32_t)(TMEM_A + 8 * (4 * pa + k)), db0 + (uint64_t)((qb * SLICE_B_BYTES + k * 32) >> 4),
                        (pa > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (g == NGROUP - 1) umma_commit(&b_empty[st]);  // the ring stage may be refilled once these MMAs have read it (and the epilogue its scales)
          umma_commit(&acc_full[g]);                       // this group's accumulators complete
        }
      }
    }
  } else {
    // =================================== epilogue ===================================
    const int q4 = warp & 3, cq = warp >> 2;   // TMEM lane quarter of this warp, column quarter of the tile
    const int row = 32 * q4 + lane;
    double sl = 0.0, sq = 0.0;
    if (cq == 0) {  // warps 0-3: this thread's row of every W' slice into tensor memory (4 K-steps of 32 bytes = 8 columns each)
#pragma unroll 1
      for (int pa = 0; pa < S; ++pa) {
        const uint4* src = reinterpret_cast<const uint4*>(wblob + (size_t)pa * SLICE_A_BYTES + (size_t)row * KP);
#pragma unroll
        for (int k = 0; k < KP / 32; ++k) {
          const uint4 u0 = src[2 * k], u1 = src[2 * k + 1];
          const uint32_t v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          tmem_st8(tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(TMEM_A + 8 * (4 * pa + k)), v);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_full);
    }
    const int rexp = (__double2hiint(reinterpret_cast<const double*>(wblob + (size_t)S * SLICE_A_BYTES)[row]) >> 20) - 1023;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int st = (int)(it % NSTAGE);
      const int64_t t0 = (tile0 + it * tstride) * NT + CPT * cq;
      // the level accumulators of this thread's 8 samples; then the accumulators go back to the MMA warp
      int32_t c[S][8];
      const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(CPT * cq);
#pragma unroll
      for (int g = 0; g < NGROUP; ++g) {
        ptx::mbar_wait(&acc_full[g], (uint32_t)(it & 1));
        tc_fence_after();
#pragma unroll
        for (int d = group_begin(g); d < group_begin(g + 1); ++d) tmem_ld8(taddr + (uint32_t)(d * NT), c[d]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[g]);
      }
      double y[CPT];
#pragma unroll
      for (int e = 0; e < CPT; ++e) {
        // sum_d C_d 128^-(d+2) = 2^-28 (hi + 2^-28 lo), hi = C0 2^14 + C1 2^7 + C2, lo = C3 2^21 + C4 2^14 + C5 2^7 + C6
        long long hi = (long long)c[0][e] * 16384 + (long long)c[1][e] * 128 + (long long)c[2][e];
        long long lo = 0;
        if (S > 3) lo = (long long)c[3][e];
#pragma unroll
        for (int d = 4; d < S; ++d) lo = lo * 128 + (long long)c[d][e];
#pragma unroll
        for (int d = S; d < 7; ++d) lo = lo * 128;
        y[e] = fma((double)lo, 3.7252902984619140625e-09 /* 2^-28 */, (double)hi);  // I2F.F64.S64: not an FP64-pipe instruction
      }
      // column scales of this thread's samples (bulk-copied with the tile: wait on its barrier for visibility; complete long ago),
      // then the ring stage goes back to the producer
      ptx::mbar_wait(&b_full[st], (uint32_t)((it / NSTAGE) & 1));
      {
        const int* csm = reinterpret_cast<const int*>(sb + (size_t)st * STAGE_BYTES + (size_t)S * SLICE_B_BYTES) + 2 * CPT * cq;
#pragma unroll
        for (int e = 0; e < CPT; ++e) y[e] = scale_pow2(y[e], rexp + (csm[2 * e + 1] >> 20) - 1023);  // scales are exact powers of two
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&b_empty[st]);
      // log-likelihood / y^2 row sums and the Y' store
      const bool partial_tile = (t0 + CPT > p.t_local);
      if (!partial_tile) {
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          double f = 0.0, fd = 0.0, dsd = 0.0;
          density_eval<DENS, false, true, BIG>(y[e], p.dp, tab, f, fd, dsd, sl);
          if (WANT_SQ) sq = fma(y[e], y[e], sq);
        }
      } else {
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          double f = 0.0, fd = 0.0, dsd = 0.0, dl = 0.0;
          density_eval<DENS, false, true, BIG>(y[e], p.dp, tab, f, fd, dsd, dl);
          if (t0 + e < p.t_local) {
            sl += dl;
            if (WANT_SQ) sq = fma(y[e], y[e], sq);
          }
        }
      }
      // Y' leaves through shared memory and the TMA unit (a thread's 8 samples are 64 bytes of ITS row: direct stores would be
      // 32 different lines per instruction).  Box = 16 samples x 128 rows, SWIZZLE_128B; the unit clips the ragged last tile
      // and the rows >= n_out.  The buffer of tile it - 2 is free: thread 0 waited for its store group before the last barrier.
      if (p.out != nullptr) {
        double* yb = ysm + (size_t)(it & 1) * 128 * NT + (size_t)(cq >> 1) * 128 * 16 + row * 16;
#pragma unroll
        for (int e = 0; e < CPT; e += 2)
          *reinterpret_cast<double2*>(yb + ((((cq & 1) * 4 + (e >> 1)) ^ (row & 7)) << 1)) = make_double2(y[e], y[e + 1]);
        ptx::fence_proxy_async();
        if (tid == 0) ptx::bulk_wait_read0();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
        if (tid == 0) {
          const int64_t tt = (tile0 + it * tstride) * NT;
          ptx::tma_store_2d(&tmap_out, ysm + (size_t)(it & 1) * 128 * NT, (int)tt, 0);
          ptx::tma_store_2d(&tmap_out, ysm + (size_t)(it & 1) * 128 * NT + 128 * 16, (int)(tt + 16), 0);
          ptx::bulk_commit();
        }
      }
    }
    if (p.out != nullptr && tid == 0) ptx::bulk_wait0();
    // the column quarters of a row live in warps q4, q4 + 4, q4 + 8, q4 + 12
    if (cq > 0) { sums[(cq - 1) * 256 + row] = sl; sums[(cq - 1) * 256 + 128 + row] = sq; }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
    if (cq == 0) {
      for (int o = 0; o < NEW / 4 - 1; ++o) { sl += sums[o * 256 + row]; sq += sums[o * 256 + 128 + row]; }
      double* out = p.partial + (size_t)blockIdx.x * rb_partial_size(128, 128, false, false);
      out[row] = 0.0; out[128 + row] = sq; out[256 + row] = sl;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace i8

int i8_mode() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("PICARD_I8");
    v = !e ? -1 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : -1));
  }
  return v;
}

size_t i8_blob_bytes(int64_t t_local) { return (size_t)((t_local + I8_TILE - 1) / I8_TILE) * I8_TILE_BYTES; }

int i8_slice_x(const double* d_x, int64_t ldx, int64_t t_local, int n, uint8_t* blob, cudaStream_t st) {
  const int64_t n_tiles = (t_local + I8_TILE - 1) / I8_TILE;
  i8::slice_x_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(d_x, ldx, t_local, n, blob);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}

template <int DENS, bool WANT_SQ>
static int launch_loss_i8_one(const PassLaunch& L, const uint8_t* xblob, uint8_t* wblob) {
  auto kern = i8::loss_i8_kernel<DENS, WANT_SQ>;
  const CUtensorMap tmap_out = L.d_out != nullptr ? make_tmap(L.d_out, L.ld_out, L.t_local, L.n_out, 128) : CUtensorMap{};
  static bool configured = false;
  if (!configured) {
    PICARD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8::SMEM_BYTES));
    configured = true;
  }
  const int64_t n_tiles = (L.t_local + I8_TILE - 1) / I8_TILE;
  int64_t grid = L.sm_count;
  if (grid > n_tiles) grid = n_tiles;
  PassParams p;
  p.w = L.d_w; p.bias = nullptr; p.n_out = L.n_out; p.n_in = L.n_in; p.ldw = L.ldw;
  p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.partial = L.d_partial; p.out = L.d_out; p.ld_out = L.ld_out;
  i8::slice_w_kernel<<<1, 1024, 0, L.stream>>>(L.d_w, L.ldw, L.n_out, L.n_in, wblob);
  PICARD_CUDA(cudaGetLastError());
  kern<<<(unsigned)grid, i8::NTHREADS, i8::SMEM_BYTES, L.stream>>>(xblob, wblob, tmap_out, p);
  PICARD_CUDA(cudaGetLastError());
  return 2 + rb_reduce(L, (int)grid, 1, 128, 128, false, false, true);
}

int launch_loss_i8(const PassLaunch& L, const uint8_t* xblob, uint8_t* wblob) {
  if (L.mode != PASS_LOSS || L.n_in > 128 || L.n_out > 128 || L.d_bias != nullptr)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the INT8 pass covers the LOSS mode for N <= 128 only");
  switch (L.dens) {
    case DENS_TANH: return L.want_h ? launch_loss_i8_one<DENS_TANH, true>(L, xblob, wblob) : launch_loss_i8_one<DENS_TANH, false>(L, xblob, wblob);
    case DENS_EXP: return L.want_h ? launch_loss_i8_one<DENS_EXP, true>(L, xblob, wblob) : launch_loss_i8_one<DENS_EXP, false>(L, xblob, wblob);
    case DENS_CUBE: return L.want_h ? launch_loss_i8_one<DENS_CUBE, true>(L, xblob, wblob) : launch_loss_i8_one<DENS_CUBE, false>(L, xblob, wblob);
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad density for the INT8 loss pass");
}

}  // namespace picard
