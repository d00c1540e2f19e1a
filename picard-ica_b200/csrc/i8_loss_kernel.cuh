// =====================================================================================================
// i8_loss_kernel.cuh -- the LOSS pass (Y' = W' x1, log-likelihood / y^2 row sums, Y' kept in HBM: core.rs:124-127 with
// compute_loss core.rs:39-85) on the INT8 tensor cores: tcgen05.mma kind::i8, level accumulators in TMEM, the operands split
// into balanced radix-256 digits (i8_common.cuh: S = 6 digits, 21 slice products, exact s32 level sums).
//
// x1 is fixed for a whole fit, so it is sliced ONCE (slice_x_kernel) into tiles of NT samples, each tile already in the shared-
// memory image the tensor core reads (K-major rows of 128 bytes = the 128 components of a sample, SWIZZLE_128B pattern), one
// [NT x 128 B] block per digit, followed by the NT sample exponents: a tile is one contiguous bulk copy (cp.async.bulk, no
// tensor map).  W' (128 x 128) is sliced by every CTA in its prologue; the partial row sums are combined, and on a single GPU the
// loss of the try computed and published to the host, in the kernel's tail (LossTail): ONE launch per line-search try.
//
// loss_i8_kernel, one CTA per SM, 18 warps:
//   warp 16 (one lane) : bulk copies of the x1 tiles into a 2-stage ring (mbarrier transaction bytes)
//   warp 17 (one lane) : per tile 21 products x 4 K-steps, several products per tcgen05.mma (kind::i8, N up to 256).  The A operand (the W'
//                        digits, 6 x 32 columns) lives in TENSOR MEMORY next to the 6 level accumulators (6 x NT columns), so
//                        the MMAs only read the B operand from shared memory; tcgen05.commit releases the ring stage and
//                        publishes the accumulators
//   warps 0-15         : epilogue, thread = (row = TMEM lane, NT / 4 samples): tcgen05.ld of the 6 levels, accumulators returned
//                        at once, exact 64-bit combination, scaling by exponent adds, log-likelihood, row sums per thread; Y'
//                        through shared memory and one TMA store per warp and tile of its own [32 rows x 16 samples] box (no
//                        CTA-wide barrier: the warps only meet at the accumulator hand-over)
// tanh log-likelihood without a logarithm per element: sum_t [|y| + log(1 + e_t) / a] = sum |y| + log(prod (1 + e_t)) / a with
// e_t = exp(-2 a |y_t|): the factors lie in [1, 2], so a thread keeps a running product, moves its exponent into an integer counter
// once per tile and takes ONE logarithm at the end of the kernel (relative error of the product ~ sqrt(#factors) 2^-53).
// ABL (ablation bits, profiles/lab only; the library instantiates ABL = 0): 1 = no density, 2 = no Y' store, 4 = trace.
// =====================================================================================================
#pragma once
#include "i8.cuh"
#include "i8_common.cuh"
#include "loss_point.cuh"
#include "p2p.cuh"

namespace picard {
namespace i8 {

constexpr int KP = 128;                  // padded contraction length = bytes per operand row
constexpr int SLICE_A_BYTES = 128 * KP;  // one digit of W': 128 rows x 128 bytes

template <int NT>
struct LossGeom {
  static_assert(NT % 16 == 0 && NT >= 16 && NT <= 64, "tile = a multiple of the UMMA N step");
  static constexpr int SLICE_B_BYTES = NT * KP;                    // one digit of a tile: NT samples x 128 bytes
  static constexpr int TILE_BYTES = S * SLICE_B_BYTES + NT * 4;    // + the NT sample exponents (int)
  static constexpr int STAGE_BYTES = ((TILE_BYTES + 1023) / 1024) * 1024;
  static constexpr int NSTAGE = 2;
  static constexpr int ACC_COLS = S * NT;                          // TMEM columns of the level accumulators of one tile
  static constexpr int TMEM_A = ACC_COLS;                          // W' digit p (p < ATM), K-step k at column TMEM_A + 8 (4 p + k)
  // W' digits kept in tensor memory: as many as fit next to the accumulators (the most significant ones: digit p takes part in
  // S - p products); the others are read from shared memory (SWIZZLE_128B image, 16 KB per digit)
  static constexpr int ATM = (512 - ACC_COLS) / 32 < S ? (512 - ACC_COLS) / 32 : S;
  static_assert(ATM >= 1 && ACC_COLS + ATM * 32 <= 512, "tensor memory: 512 columns");
  static constexpr int NEW = 16;                                   // epilogue warps: TMEM lane quarter (warp & 3) x column quarter (warp >> 2)
  static constexpr int CPT = NT / 4;                               // samples per epilogue thread and tile
  static_assert(CPT % 4 == 0, "tcgen05.ld x4 / x8 granularity");
  static constexpr int NTHREADS = 32 * (NEW + 2);
  static_assert(CPT == 16, "the Y' box row is one 128-byte swizzle row");
  static constexpr size_t SMEM_Y = (size_t)NEW * 32 * CPT * 8;      // one [32 rows x CPT samples] TMA-store box per epilogue warp
  static constexpr size_t SMEM_B = (size_t)NSTAGE * STAGE_BYTES;
  static constexpr size_t SMEM_A = (size_t)(S - ATM) * SLICE_A_BYTES;
  static constexpr bool BIG = true;                                // the 2048-entry exp table (16 KB); the log table is not needed
  static constexpr size_t TAB_BYTES = (size_t)dmath::Tab<BIG>::EXP_N * 8;
  static constexpr size_t SMEM_BYTES = SMEM_Y + SMEM_B + SMEM_A + TAB_BYTES + 8192 /* sums */ + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static constexpr uint32_t IDESC = make_idesc(NT);
};

// What follows the streaming part inside the kernel (instead of two more launches per line-search try): the per-CTA partial row sums
// are combined in a fixed order by CTA 0 once every CTA has published its partial (device-wide counter), and -- on a single GPU,
// where no exchange between ranks comes in between -- the loss of the try and its accept flag are computed and published to the
// host (core.rs:39-85 after the sums, core.rs:127-132).
struct LossTail {
  unsigned int* counter = nullptr;   // device, monotonically increasing over launches; nullptr: no tail (partials only)
  unsigned int target = 0;           // counter value once every CTA of THIS launch has arrived
  double* mom = nullptr;             // reduced moments: the Sq and L sections are written
  int finish = 0;                    // 1: also loss + accept flag + publish (single GPU)
  int which = 0;                     // 0: a line-search try (new_loss, accept) ; 1: the current point (current_loss, loss_singular)
  CoreDims dims{};
  const double* signs = nullptr;
  CoreScalars* sc = nullptr;
  CoreScalars* sc_map = nullptr;
  unsigned long long seq = 0;
  // several GPUs: the 2 n reduced row sums are exchanged over peer memory in this tail too (one-shot: pushed to every rank's
  // mailbox, summed in rank order on every rank), so a line-search try is ONE launch on any number of GPUs
  int exchange = 0;
  P2PCall px{};
};

// ---------------------------------------------------------------------------------------------------
// slicing of x1: a CTA per tile of NT samples (grid-stride), thread = (sample r, 16-byte chunk ch of its 128 digit bytes).
// Also gathers the statistics the range guard and the gradient pass need (I8_XSTATS).
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(8 * NT) slice_x_kernel(const double* __restrict__ x, int64_t ldx, int64_t t_local, int n, int64_t n_tiles,
                                                         uint8_t* __restrict__ blob, double* __restrict__ stats) {
  using G = LossGeom<NT>;
  __shared__ double red[8 * NT / 32][128];
  const int tid = threadIdx.x, r = tid >> 3, ch = tid & 7, warp = tid >> 5, lane = tid & 31;
  double rs[16];
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) rs[kk] = 0.0;
  double sum_scale = 0.0, max_n2 = 0.0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t t = tile * NT + r;
    double v[16];
    double m = 0.0, n2 = 0.0;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = 16 * ch + kk;
      v[kk] = (k < n && t < t_local) ? x[(size_t)k * ldx + t] : 0.0;
      m = fmax(m, fabs(v[kk]));
      n2 = fma(v[kk], v[kk], n2);
      rs[kk] = fma(v[kk], v[kk], rs[kk]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    const int e = bound_exponent(m);
    if (ch == 0 && t < t_local) {
      sum_scale += scalbn(1.0, e - 1);
      max_n2 = fmax(max_n2, n2);
    }
    uint64_t dg[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) dg[kk] = split_digits(v[kk], e);
    uint8_t* tb = blob + (size_t)tile * G::TILE_BYTES;
#pragma unroll
    for (int p = 0; p < S; ++p) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) w[kk >> 2] |= (uint32_t)((dg[kk] >> (8 * (S - 1 - p))) & 0xFF) << (8 * (kk & 3));
      *reinterpret_cast<uint4*>(tb + (size_t)p * G::SLICE_B_BYTES + r * KP + ((ch ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (ch == 0) reinterpret_cast<int*>(tb + (size_t)S * G::SLICE_B_BYTES)[r] = e;
  }
  // statistics: row sums of squares over this CTA's samples (threads of equal ch hold the same 16 rows), scale sum, max norm
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    rs[kk] += __shfl_xor_sync(0xffffffffu, rs[kk], 8);
    rs[kk] += __shfl_xor_sync(0xffffffffu, rs[kk], 16);
  }
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {
    sum_scale += __shfl_xor_sync(0xffffffffu, sum_scale, o);
    max_n2 = fmax(max_n2, __shfl_xor_sync(0xffffffffu, max_n2, o));
  }
  if (lane < 8) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) red[warp][16 * lane + kk] = rs[kk];
  }
  if (lane == 0) {
    atomicAdd(&stats[0], sum_scale);
    atomicMax(reinterpret_cast<unsigned long long*>(&stats[1]), (unsigned long long)__double_as_longlong(max_n2));
  }
  __syncthreads();
  if (tid < 128) {
    double s = 0.0;
    for (int w = 0; w < 8 * NT / 32; ++w) s += red[w][tid];
    atomicAdd(&stats[2 + tid], s);
  }
}

#ifndef I8_TRACE_SLOTS
#define I8_TRACE_SLOTS 0
#endif

// ---------------------------------------------------------------------------------------------------
// the pass
// ---------------------------------------------------------------------------------------------------
template <int DENS, bool WANT_SQ, int NT, int ABL = 0>
__global__ void __launch_bounds__(LossGeom<NT>::NTHREADS, 1)  // 18 warps are allocated as 20: 96 registers per thread
loss_i8_kernel(const uint8_t* __restrict__ xblob, const __grid_constant__ CUtensorMap tmap_out, const PassParams p, const LossTail tail,
               long long* __restrict__ trace) {
  using G = LossGeom<NT>;
  constexpr int NEW = G::NEW, CPT = G::CPT, NSTAGE = G::NSTAGE;
  constexpr bool BIG = G::BIG;
  constexpr bool NO_DENS = (ABL & 1) != 0, NO_STORE = (ABL & 2) != 0, TRACE = (ABL & 4) != 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  double* ysm = reinterpret_cast<double*>(smem);
  unsigned char* sb = smem + G::SMEM_Y;
  unsigned char* sa = smem + G::SMEM_Y + G::SMEM_B;   // W' digits ATM .. S-1 (SWIZZLE_128B image), 1024-byte aligned
  double* tab = reinterpret_cast<double*>(smem + G::SMEM_Y + G::SMEM_B + G::SMEM_A);
  double* sums = tab + dmath::Tab<BIG>::EXP_N;  // [NEW / 4 - 1][2][128]: partial row sums of the column quarters
  uint64_t* bars = reinterpret_cast<uint64_t*>(sums + 1024);
  uint64_t* a_full = bars;                 // W' digits stored in tensor memory (4 warps)
  uint64_t* b_full = bars + 1;             // [NSTAGE] tile landed (transaction bytes)
  uint64_t* b_empty = b_full + NSTAGE;     // [NSTAGE] MMA commit + the NEW epilogue warps (they read the sample exponents)
  uint64_t* acc_full = b_empty + NSTAGE;   // accumulators of a tile complete (MMA commit)
  uint64_t* acc_empty = acc_full + 1;      // accumulators read back (NEW epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool NEED_TAB = !NO_DENS && (DENS == DENS_TANH || DENS == DENS_EXP);
  if (NEED_TAB) load_density_tables<BIG>(tab, false, tid, G::NTHREADS);
  if (tid == 0) {
    ptx::mbar_init(a_full, NEW);
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1 + NEW); }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, NEW);
    ptx::fence_barrier_init();
  }
  if (warp == NEW + 1) tmem_alloc512(tmem_slot);  // the MMA warp owns the tensor memory: all 512 columns (one CTA per SM)
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
        ptx::mbar_expect_tx(&b_full[st], (uint32_t)G::TILE_BYTES);
        bulk_load(sb + (size_t)st * G::STAGE_BYTES, xblob + (size_t)(tile0 + it * tstride) * G::TILE_BYTES, G::TILE_BYTES, &b_full[st]);
      }
    }
  } else if (warp == NEW + 1) {
    // =================================== MMA issue ===================================
    if (lane == 0) {
      ptx::mbar_wait(a_full, 0);
      tc_fence_after();
      const uint64_t da0 = make_desc(smem_u32(sa));
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % NSTAGE);
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 0] = clock64();
        ptx::mbar_wait(&b_full[st], (uint32_t)((it / NSTAGE) & 1));
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 1] = clock64();
        const uint64_t db0 = make_desc(smem_u32(sb + (size_t)st * G::STAGE_BYTES));
        ptx::mbar_wait(acc_empty, (uint32_t)((it & 1) ^ 1));  // the epilogue has read tile it - 1 back (first tile: passes)
        tc_fence_after();
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 2] = clock64();
        // One MMA covers several digit products at once: the digits of a tile are adjacent blocks of NT rows in shared memory and
        // the level accumulators adjacent blocks of NT columns in tensor memory, so W' digit p against x digits q .. q + m - 1 is ONE
        // 128 x (m NT) x 32 MMA writing levels p + q .. p + q + m - 1 (N up to 256: the INT8 tensor rate needs N >= 128,
        // profiles/microbench/pipe_probe_r02.jsonl).  Digit 0 of W' goes first: it touches every level, so its MMAs of the first
        // K-step initialise all accumulators and everything else accumulates.
        constexpr int MAXD = 256 / NT;  // digits per MMA
#pragma unroll
        for (int k = 0; k < KP / 32; ++k) {
#pragma unroll
          for (int pa = 0; pa < S; ++pa) {
#pragma unroll
            for (int qb = 0; qb < S - pa; qb += MAXD) {
              const int m = (S - pa - qb) < MAXD ? (S - pa - qb) : MAXD;
              const uint32_t tacc = tmem + (uint32_t)((pa + qb) * NT);
              const uint64_t db = db0 + (uint64_t)((qb * G::SLICE_B_BYTES + k * 32) >> 4);
              const uint32_t acc = (pa > 0 || k > 0) ? 1u : 0u;
              if ((ABL & 8) != 0) continue;  // lab ablation: the epilogue alone, on whatever the accumulators hold
              if (pa < G::ATM) umma_i8_ts(tacc, tmem + (uint32_t)(G::TMEM_A + 8 * (4 * pa + k)), db, make_idesc(m * NT), acc);
              else umma_i8_ss(tacc, da0 + (uint64_t)(((pa - G::ATM) * SLICE_A_BYTES + k * 32) >> 4), db, make_idesc(m * NT), acc);
            }
          }
        }
        umma_commit(&b_empty[st]);  // the ring stage may be refilled once these MMAs have read it (and the epilogue its exponents)
        umma_commit(acc_full);      // the accumulators are complete
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 3] = clock64();
      }
    }
  } else {
    // =================================== epilogue ===================================
    const int q4 = warp & 3, cq = warp >> 2;   // TMEM lane quarter of this warp, column quarter of the tile
    const int row = 32 * q4 + lane;
    double sl = 0.0, sq = 0.0, prod = 1.0;
    int pexp = 0;
    // ---- W' (n_out x n_in f64, zero-padded to 128 x 128) -> balanced digits, in the kernel (no separate slicing launch per try):
    // thread = (row, K-step cq = 32 of its 128 columns).  Row maximum across the four K-step warps through shared memory, then
    // the digits of the thread's 32 values: one 32-byte K-step row per digit, into tensor memory (digits < ATM, tcgen05.st) or
    // into the SWIZZLE_128B image in shared memory (the least significant digits).
    int rexp;
    {
      const double* wrow = p.w + (size_t)row * p.ldw + 32 * cq;
      double m = 0.0;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const double v = (row < p.n_out && 32 * cq + k < p.n_in) ? wrow[k] : 0.0;
        m = fmax(m, fabs(v));
      }
      sums[cq * 128 + row] = m;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
      m = fmax(fmax(sums[row], sums[128 + row]), fmax(sums[256 + row], sums[384 + row]));
      const int e = bound_exponent(m);
      rexp = e + COMBINE_EXP;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint64_t dg[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int col = 32 * cq + 16 * half + k;
          dg[k] = split_digits((row < p.n_out && col < p.n_in) ? wrow[16 * half + k] : 0.0, e);
        }
#pragma unroll
        for (int pa = 0; pa < S; ++pa) {
          uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int k = 0; k < 16; ++k) w4[k >> 2] |= (uint32_t)((dg[k] >> (8 * (S - 1 - pa))) & 0xFF) << (8 * (k & 3));
          if (pa < G::ATM) tmem_st4(tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(G::TMEM_A + 8 * (4 * pa + cq) + 4 * half), w4);
          else *reinterpret_cast<uint4*>(sa + (size_t)(pa - G::ATM) * SLICE_A_BYTES + row * KP + (((2 * cq + half) ^ (row & 7)) << 4)) =
                   make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
      tmem_wait_st();
      ptx::fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");  // `sums` is reused for the row sums at the end
    }
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int st = (int)(it % NSTAGE);
      const int64_t t0 = (tile0 + it * tstride) * NT + CPT * cq;
      const bool tr = TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS && warp == 0 && lane == 0;
      if (tr) trace[it * 8 + 4] = clock64();
      // the level accumulators of this thread's samples, read in two groups of three levels that are combined into 64-bit integers at
      // once (CPT x 6 raw levels would not fit the 96 registers of a thread at CPT = 16); then the accumulators go back to the MMA warp
      const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(CPT * cq);
      ptx::mbar_wait(acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (tr) trace[it * 8 + 5] = clock64();
      double y[CPT];
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        int32_t c[3][CPT];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const uint32_t ta = taddr + (uint32_t)((3 * grp + d) * NT);
          if (CPT == 16) tmem_ld16(ta, &c[d][0]);
          else {
#pragma unroll
            for (int o = 0; o + 8 <= CPT; o += 8) tmem_ld8(ta + (uint32_t)o, &c[d][o]);
            if (CPT % 8 == 4) tmem_ld4(ta + (uint32_t)(CPT - 4), &c[d][CPT - 4]);
          }
        }
        tmem_wait_ld();
        if (grp == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        // = combine_levels(): hi + 2^-24 lo, rounded once -- without the two 64-bit integer-to-double conversions per element (XU
        // pipe, 8 cycles per warp instruction): for |v| < 2^51 the bit pattern of M + v (M = 1.5 2^52) is M's pattern plus v, so
        //   hi + 2^-24 lo = fma(M + lo, 2^-24, (M + hi) - M (1 + 2^-24)),  every intermediate exact, one rounding at the end
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          const long long v = (long long)c[0][e] * 65536 + (long long)c[1][e] * 256 + (long long)c[2][e];
          const double raw = __longlong_as_double(v + 0x4338000000000000LL);  // M + v
          if (grp == 0) y[e] = raw - 6755399441055744.0 * (1.0 + 5.9604644775390625e-08);
          else y[e] = fma(raw, 5.9604644775390625e-08 /* 2^-24 */, y[e]);
        }
      }
      if (tr) trace[it * 8 + 6] = clock64();
      // exponents of this thread's samples (bulk-copied with the tile: wait on its barrier for visibility; complete long ago)
      // -- the ring stage goes back to the producer further down, once they have been used
      ptx::mbar_wait(&b_full[st], (uint32_t)((it / NSTAGE) & 1));
      {
        const int* csm = reinterpret_cast<const int*>(sb + (size_t)st * G::STAGE_BYTES + (size_t)S * G::SLICE_B_BYTES) + CPT * cq;
#pragma unroll
        for (int e = 0; e < CPT; ++e) y[e] = scale_pow2(y[e], rexp + csm[e]);
      }
      // Y' leaves through shared memory and ONE TMA store per warp and tile: box = this warp's [32 rows x CPT samples] (SWIZZLE_128B:
      // 16-byte chunk c of row r at position c ^ (r & 7), conflict-free for lane = row; the unit clips the ragged last tile and the
      // rows >= n_out).  The values are written NOW and the store is issued after the density arithmetic below, so that the proxy
      // fence between the two does not wait for the shared-memory writes (issued right after them it cost 9 % of the epilogue).
      if (!NO_STORE && p.out != nullptr) {
        unsigned char* yb = reinterpret_cast<unsigned char*>(ysm) + (size_t)warp * (32 * CPT * 8);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the unit has read the previous tile's box
        __syncwarp();
#pragma unroll
        for (int c2 = 0; c2 < CPT / 2; ++c2)
          *reinterpret_cast<double2*>(yb + lane * (CPT * 8) + ((c2 ^ (lane & (CPT / 2 - 1))) << 4)) = make_double2(y[2 * c2], y[2 * c2 + 1]);
      } else {
        // no Y' store: one shared-memory write that depends on every scaled value, so that the release below still follows the use of
        // everything read from the stage
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < CPT; ++e) acc += y[e];
        reinterpret_cast<volatile double*>(reinterpret_cast<unsigned char*>(ysm) + (size_t)warp * (32 * CPT * 8))[lane] = acc;
      }
      // The ring stage goes back to the producer only after the shared-memory writes that consume the exponents read from it: an
      // arrival issued right behind the loads can overtake them while they are in flight (found in the gradient kernel, where the
      // refill then replaced rows under the loads; see the note there).
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[st]);
      // log-likelihood / y^2 row sums
      if (!NO_DENS) {
        const bool partial_tile = (t0 + CPT > p.t_local);
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          const bool valid = !partial_tile || (t0 + e < p.t_local);  // loglik(0) != 0: padding columns must not reach L
          if (DENS == DENS_TANH) {
            // |y| on the ALU, clamped: every FP64-pipe instruction here competes with the tensor core's INT8 MMAs for the pipe
            // (profiles/microbench/pipe_probe_r02.jsonl)
            const int hy = __double2hiint(y[e]) & 0x7fffffff;
            const double ayc = __hiloint2double(hy < p.dp.hi_limit ? hy : p.dp.hi_limit, __double2loint(y[e]));
            const double ez = dmath::exp_scaled<BIG>(ayc, p.dp, tab);  // exp(-2 alpha |y|)
            sl += fabs(y[e]);  // |0| = 0: no mask needed (the absolute value is an operand modifier of the add)
            prod *= (partial_tile && !valid) ? 1.0 : 1.0 + ez;
          } else {
            double f = 0.0, fd = 0.0, dsd = 0.0, dl = 0.0;
            density_eval<DENS, false, true, BIG>(y[e], p.dp, tab, f, fd, dsd, dl);
            if (valid) sl += dl;
          }
          if (WANT_SQ) sq = fma(y[e], y[e], sq);
        }
        if (DENS == DENS_TANH) {  // prod in [1, 2^(CPT+1)): its exponent goes to the counter, the mantissa stays in [1, 2)
          const int h = __double2hiint(prod), k = (h >> 20) - 1023;
          pexp += k;
          prod = __hiloint2double(h - (k << 20), __double2loint(prod));
        }
      } else {
#pragma unroll
        for (int e = 0; e < CPT; ++e) sl += y[e];
      }
      // ... and the TMA store of this warp's box, now that its shared-memory writes (issued before the density arithmetic) have long
      // completed: the proxy fence finds nothing to wait for
      if (!NO_STORE && p.out != nullptr) {
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_out, ysm + (size_t)warp * (32 * CPT), (int)((tile0 + it * tstride) * NT + CPT * cq), 32 * q4);
          bulk_commit();
        }
      }
      if (tr) trace[it * 8 + 7] = clock64();
    }
    if (!NO_STORE && p.out != nullptr && lane == 0) bulk_wait0();
    if (DENS == DENS_TANH && !NO_DENS) sl = fma(fma((double)pexp, 6.931471805599453094e-01, log(prod)), p.dp.inv_alpha, sl);
    // the column quarters of a row live in warps q4, q4 + 4, q4 + 8, q4 + 12
    if (cq > 0) { sums[(cq - 1) * 256 + row] = sl; sums[(cq - 1) * 256 + 128 + row] = sq; }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
    if (cq == 0) {
      for (int o = 0; o < NEW / 4 - 1; ++o) { sl += sums[o * 256 + row]; sq += sums[o * 256 + 128 + row]; }
      double* out = p.partial + (size_t)blockIdx.x * (3 * 128);  // [Sd (unused) | Sq | L] of the 128 rows: rb_partial_size(128, 128, false, false)
      out[row] = 0.0; out[128 + row] = sq; out[256 + row] = sl;
    }
  }
  __threadfence();  // the partial (written by the cq == 0 threads above) is visible device-wide before this CTA is counted
  tc_fence_before();
  __syncthreads();
  if (warp == NEW + 1) tmem_dealloc512(tmem);
  if (tail.counter != nullptr) {
    if (blockIdx.x != 0) {
      if (tid == 0) atomicAdd(tail.counter, 1u);
    } else {
      if (tid == 0) {
        atomicAdd(tail.counter, 1u);
        while ((int)(*reinterpret_cast<volatile unsigned int*>(tail.counter) - tail.target) < 0) __nanosleep(64);
        __threadfence();
      }
      __syncthreads();
      // fixed-order sum of the per-CTA partials [Sd (unused) | Sq | L]: eight independent chains per element (latency-bound)
      if (tid < 256) {
        const int r = tid & 127, sec = tid >> 7;
        const double* pp = p.partial + 128 + sec * 128 + r;
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int b = 0;
        for (; b + 8 <= (int)gridDim.x; b += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] += __ldcg(pp + (size_t)(b + u) * 384);
        }
        for (int u = 0; b < (int)gridDim.x; ++b, ++u) acc[u] += __ldcg(pp + (size_t)b * 384);
        if (r < p.n_out)
          tail.mom[(sec == 0 ? mom_off_sq(p.n_out) : mom_off_ll(p.n_out)) + r] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
      }
      __syncthreads();
      if (tail.exchange) {
        const int n = p.n_out;
        double* seg = tail.mom + mom_off_sq(n);  // [Sq n][L n], contiguous
        for (int q = 0; q < tail.px.nranks; ++q) {
          double* slot = tail.px.peers.box[q] + p2p_slot(tail.px.parity, tail.px.rank);
          for (int i = tid; i < 2 * n; i += G::NTHREADS) slot[i] = seg[i];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < tail.px.nranks) {
          atomicAdd_system(tail.px.peers.flags[tid] + tail.px.parity * P2P_MAX_RANKS + tail.px.rank, (unsigned int)P2P_PARTS);
          volatile unsigned int* f = tail.px.peers.flags[tail.px.rank] + tail.px.parity * P2P_MAX_RANKS + tid;
          while ((int)(*f - tail.px.expect) < 0) __nanosleep(32);
          __threadfence_system();
        }
        __syncthreads();
        const double* box = tail.px.peers.box[tail.px.rank] + p2p_slot(tail.px.parity, 0);
        for (int i = tid; i < 2 * n; i += G::NTHREADS) {
          double acc = __ldcv(box + i);
          for (int q = 1; q < tail.px.nranks; ++q) acc += __ldcv(box + (size_t)q * P2P_MAX_DOUBLES + i);
          seg[i] = acc;
        }
        __syncthreads();
      }
      if (tail.finish && warp == 0) {
        bool sing;
        const double l = small::loss_of_point_warp(tail.dims, tail.mom, tail.signs, &sing);
        if (lane == 0) {
          if (tail.which == 0) {
            tail.sc->new_loss = l;
            tail.sc->accept = (l < tail.sc->current_loss) ? 1 : 0;
          } else {
            tail.sc->current_loss = l;
            tail.sc->loss_singular = sing ? 1 : 0;
          }
          publish_scalars(tail.sc, tail.sc_map, tail.seq);
        }
      }
    }
  }
}

}  // namespace i8
}  // namespace picard
