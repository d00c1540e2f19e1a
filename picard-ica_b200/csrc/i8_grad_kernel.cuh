// =====================================================================================================
// i8_grad_kernel.cuh -- the stored-Y gradient pass Gr = psi(Y) Y^T, Sd = sum_t psi'(Y) (core.rs:215-218, 226) on the INT8
// tensor cores.  The contraction runs over SAMPLES, so both operands are new every iteration and are split into balanced
// radix-256 digits (i8_common.cuh) inside the kernel, by the same warps that evaluate psi.
//
// Scales.  psi is bounded by a constant of the density (|tanh| <= 1, |y exp(-a y^2 / 2)| <= 1/sqrt(a e)): one fixed exponent.
// Row j of Y gets ONE exponent for the whole shard from the rigorous bound |y_jt| <= |w_j|_2 max_t |x_t|_2 (the row norms of the W
// that produced Y, the largest sample norm of x1 gathered once per fit by slice_x_kernel): no pass over Y is needed to find it.
// With fixed exponents the level sums of all samples of a CTA add up exactly in the s32 accumulators; they are flushed to f64
// every FLUSH_TILES tiles (bound: 6 products x 16384 samples x 128^2 < 2^31).
//
// Work split.  D[j][i] = sum_t y_jt psi_it with M = 128 rows j (A operand: the digits of ALL rows of Y) and N = 64 columns i (B
// operand: the digits of psi of the CTA's own 64 rows): 6 levels x 64 columns = 384 of the 512 TMEM columns.  Two CTAs (the two
// halves of i) share a tile group and read the same Y tiles (the second read hits L2); each evaluates psi only for its own rows,
// so no density work is duplicated -- only the cheap digit conversion of Y (8 instructions / element).
//
// One CTA per SM, 18 warps:
//   warp 16 (one lane) : TMA loads of the Y tiles ([128 rows x 32 samples] f64 as two SWIZZLE_128B boxes) into a YS-stage ring
//   warp 17 (one lane) : per tile 21 tcgen05.mma (128 x 64 x 32, kind::i8, both operands from shared memory), commit -> slot free
//   warps 0-15         : converters.  Every thread converts 8 samples of one row of Y (y -> fixed point by one FMA with magic-number
//                        rounding, digits by one 64-bit add + byte permutes, STS.64 into the operand slot: K-major rows of 32
//                        bytes, SWIZZLE_32B, one UMMA K-step per row) AND evaluates psi for 4 samples of one of the CTA's own 64
//                        rows (the same work for every thread: the four warps of a scheduler stay balanced), converts it and
//                        stores it (STS.32).  Every FLUSH_TILES tiles the warps read the level accumulators back (tcgen05.ld),
//                        combine them exactly and add them to the CTA's f64 partial of Gr.
// FP64-pipe economy (the INT8 MMAs and FP64 instructions share a pipe: profiles/microbench/pipe_probe_r02.jsonl): |y|, clamps and
// sign transfers on the ALU; tanh: Sd = alpha (T - sum tanh^2), one FMA per element instead of psi' and its sum.
// =====================================================================================================
#pragma once
#include "i8.cuh"
#include "i8_common.cuh"
#include "p2p.cuh"

namespace picard {
namespace i8 {

struct GradGeom {
  static constexpr int KT = 32;                    // samples per tile = one UMMA K-step of int8
  static constexpr int MA = 128;                   // rows of the A operand (all rows of Y)
  static constexpr int NB = 64;                    // rows of the B operand (psi of the CTA's own rows) = accumulator columns per level
  static constexpr int YS = 3;                     // stages of the Y ring
  static constexpr int SLOTS = 3;                  // operand slots
  static constexpr int Y_STAGE_BYTES = MA * KT * 8;          // 32768: two [128 x 16] f64 boxes
  static constexpr int A_DIGIT_BYTES = MA * KT;              // 4096
  static constexpr int B_DIGIT_BYTES = NB * KT;              // 2048
  static constexpr int SLOT_BYTES = S * (A_DIGIT_BYTES + B_DIGIT_BYTES);  // 36864
  static constexpr int NCW = 16;                   // converter warps
  static constexpr int NTHREADS = 32 * (NCW + 2);
  static constexpr int FLUSH_TILES = 512;          // 16384 samples between flushes
  static constexpr size_t TAB_BYTES = (size_t)dmath::Tab<true>::EXP_N * 8;  // the exp table only (no log-likelihood here)
  static constexpr size_t TAIL_BYTES = 3072;  // barriers (14 x 8), the TMEM slot, [4][64] Sd partials
  static constexpr size_t SMEM_BYTES = (size_t)YS * Y_STAGE_BYTES + (size_t)SLOTS * SLOT_BYTES + TAB_BYTES + TAIL_BYTES;
  static constexpr uint32_t IDESC = make_idesc(NB);
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert((long long)S * FLUSH_TILES * KT * 128 * 128 < (1ll << 31), "level sums must stay exact in s32");
};

// Operand slot layouts (K-major, one 32-byte K-step per row; 8-row groups 256 bytes apart):
//   LAYOUT 0: SWIZZLE_32B -- rows of 32 bytes, 16-byte chunk c of row r at chunk position c ^ ((r >> 2) & 1)
//   LAYOUT 1: no swizzle (interleaved core matrices of 8 rows x 16 bytes) -- chunk c of row r at (r >> 3) 256 + c 128 + (r & 7) 16
// Both are conflict-free for the converters' STS.64 and both are checked on the hardware by profiles/lab; the library uses LAYOUT 1
// (measured 5.57 against 5.78 ms at c3).
template <int LAYOUT>
__device__ __forceinline__ uint64_t make_desc_k32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LAYOUT == 0 ? 1 : (128 >> 4)) << 16;  // LBO: K-adjacent core matrices (unused for swizzled K-major layouts)
  d |= (uint64_t)(256 >> 4) << 32;                      // SBO: 8-row groups
  d |= (uint64_t)1 << 46;                               // descriptor version of sm_100
  d |= (uint64_t)(LAYOUT == 0 ? 6 : 0) << 61;           // SWIZZLE_32B / SWIZZLE_NONE
  return d;
}
template <int LAYOUT>
__device__ __forceinline__ uint32_t slot_row_offset(int row, int chunk) {
  return LAYOUT == 0 ? (uint32_t)(row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4)) : (uint32_t)((row >> 3) * 256 + chunk * 128 + (row & 7) * 16);
}

// 8 values v[i] (|v[i] sc| <= 2^46) -> their six balanced digits, packed per digit: w[p][h] holds byte p (p = 0 least significant,
// i.e. digit S - 1 - p) of samples 4h .. 4h+3.  One FMA per value (magic-number rounding to the nearest integer), one 64-bit add,
// 12 byte permutes and 6 XORs per four values.
__device__ __forceinline__ void digits8(const double (&v)[8], double sc, uint32_t (&w)[S][2]) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double u = fma(v[4 * h + i], sc, MAGIC);  // mantissa field = 2^51 + I
      const unsigned long long U = (unsigned long long)__double_as_longlong(u) + DIGIT_BIAS;
      lo[i] = (uint32_t)U; hi[i] = (uint32_t)(U >> 32);
    }
    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[0], lo[1], 0x7362);
    const uint32_t u0 = __byte_perm(lo[2], lo[3], 0x5140), u1 = __byte_perm(lo[2], lo[3], 0x7362);
    const uint32_t v0 = __byte_perm(hi[0], hi[1], 0x5140), v1 = __byte_perm(hi[2], hi[3], 0x5140);
    w[0][h] = __byte_perm(t0, u0, 0x5410) ^ 0x80808080u;
    w[1][h] = __byte_perm(t0, u0, 0x7632) ^ 0x80808080u;
    w[2][h] = __byte_perm(t1, u1, 0x5410) ^ 0x80808080u;
    w[3][h] = __byte_perm(t1, u1, 0x7632) ^ 0x80808080u;
    w[4][h] = __byte_perm(v0, v1, 0x5410) ^ 0x80808080u;
    w[5][h] = __byte_perm(v0, v1, 0x7632) ^ 0x80808080u;
  }
}

// four values -> six words (word p = byte p of the four fixed-point integers)
__device__ __forceinline__ void digits4(const double (&v)[4], double sc, uint32_t (&w)[S]) {
  const double MAGIC = 6755399441055744.0;
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double u = fma(v[i], sc, MAGIC);
    const unsigned long long U = (unsigned long long)__double_as_longlong(u) + DIGIT_BIAS;
    lo[i] = (uint32_t)U; hi[i] = (uint32_t)(U >> 32);
  }
  const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[0], lo[1], 0x7362);
  const uint32_t u0 = __byte_perm(lo[2], lo[3], 0x5140), u1 = __byte_perm(lo[2], lo[3], 0x7362);
  const uint32_t v0 = __byte_perm(hi[0], hi[1], 0x5140), v1 = __byte_perm(hi[2], hi[3], 0x5140);
  w[0] = __byte_perm(t0, u0, 0x5410) ^ 0x80808080u;
  w[1] = __byte_perm(t0, u0, 0x7632) ^ 0x80808080u;
  w[2] = __byte_perm(t1, u1, 0x5410) ^ 0x80808080u;
  w[3] = __byte_perm(t1, u1, 0x7632) ^ 0x80808080u;
  w[4] = __byte_perm(v0, v1, 0x5410) ^ 0x80808080u;
  w[5] = __byte_perm(v0, v1, 0x7632) ^ 0x80808080u;
}

// tanh(alpha y) with |.|, clamp and sign transfer on the ALU: 13 FP64-pipe instructions
__device__ __forceinline__ double tanh_psi(double y, const DensParams& dp, const double* __restrict__ tab) {
  const int hy = __double2hiint(y), ha = hy & 0x7fffffff;
  const double ayc = __hiloint2double(ha < dp.hi_limit ? ha : dp.hi_limit, __double2loint(y));
  const double e = dmath::exp_scaled<true>(ayc, dp, tab);  // exp(-2 alpha |y|)
  const double th = dmath::div_seeded(1.0 - e, 1.0 + e);   // tanh(alpha |y|) >= 0
  return __hiloint2double(__double2hiint(th) | (hy & 0x80000000), __double2loint(th));
}

// psi exponent: 1.008 |psi| < 2^e for every argument
inline int psi_exponent(int dens, double alpha) {
  if (dens == DENS_TANH) return 1;                                     // |tanh| <= 1: I <= 2^46
  return bound_exponent(1.0000001 / std::sqrt(alpha * 2.718281828459045));  // exp: max |y exp(-a y^2/2)| = 1/sqrt(a e)
}

struct GradParams {
  int n;                 // rows of Y (<= 128)
  int64_t t_local, n_tiles;
  DensParams dp;
  const int* rowexp;     // n ints: e_j with 1.008 |y_jt| < 2^e_j
  int psi_exp;
  double* partial;       // [gridDim.x][rb_partial_size(64, 128, true, false)]
  // tail (optional): every CTA waits until all partials are published (device-wide counter; the CTAs are co-resident: cooperative
  // launch, one CTA per SM) and then sums its slice of [Gr | Sd] over the tile groups in a fixed order into the moment buffer
  unsigned int* counter; // nullptr: no tail (the host launches the reduction)
  unsigned int target;
  double* mom;
  // several GPUs: the exchange of [Gr | Sd] between the ranks happens in the tail as well (compute + reduce + allreduce in ONE kernel):
  // every CTA pushes its reduced slice into every rank's mailbox; the CTA that completes the local push signals the peers; once
  // every rank's contribution has arrived each CTA sums its slice over the ranks in rank order.
  int exchange;
  unsigned int* counter2; unsigned int target2;   // second device-wide count: "every CTA has pushed"
  P2PCall px;
};

#ifndef I8_TRACE_SLOTS
#define I8_TRACE_SLOTS 0
#endif

template <int DENS, int ABL = 0, int LAYOUT = 0>
__global__ void __launch_bounds__(GradGeom::NTHREADS, 1)
grad_i8_kernel(const __grid_constant__ CUtensorMap tmap, const GradParams p, long long* __restrict__ trace) {
  using G = GradGeom;
  constexpr bool NO_PSI = (ABL & 1) != 0, TRACE = (ABL & 4) != 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ysm = smem;                                            // [YS][2 boxes][128 rows][16 samples] f64, SWIZZLE_128B
  unsigned char* osm = smem + (size_t)G::YS * G::Y_STAGE_BYTES;         // [SLOTS][A: S x 4096 | B: S x 2048]
  double* tab = reinterpret_cast<double*>(osm + (size_t)G::SLOTS * G::SLOT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tab) + G::TAB_BYTES);
  uint64_t* y_full = bars;                        // [YS] tile landed (transaction bytes)
  uint64_t* y_empty = y_full + G::YS;             // [YS] converters done with the stage (NCW warps)
  uint64_t* o_full = y_empty + G::YS;             // [SLOTS] digits written (NCW warps)
  uint64_t* o_empty = o_full + G::SLOTS;          // [SLOTS] MMAs of the slot complete (commit)
  uint64_t* f_full = o_empty + G::SLOTS;          // accumulators complete up to a flush point (commit)
  uint64_t* f_empty = f_full + 1;                 // accumulators read back (NCW warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(f_empty + 1);
  double* sdsm = reinterpret_cast<double*>(tmem_slot + 2);  // [4][64]: Sd partials of the (sample half, sample quarter) warp groups
  static_assert((2 * G::YS + 2 * G::SLOTS + 2) * 8 + 8 + 4 * 64 * 8 <= G::TAIL_BYTES, "shared-memory tail");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.x & 1, tg = blockIdx.x >> 1, n_tg = gridDim.x >> 1;
  constexpr bool NEED_TAB = !NO_PSI && (DENS == DENS_TANH || DENS == DENS_EXP);
  if (NEED_TAB) load_density_tables<true>(tab, false, tid, G::NTHREADS);
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int s = 0; s < G::YS; ++s) { ptx::mbar_init(&y_full[s], 1); ptx::mbar_init(&y_empty[s], G::NCW); }
    for (int s = 0; s < G::SLOTS; ++s) { ptx::mbar_init(&o_full[s], G::NCW); ptx::mbar_init(&o_empty[s], 1); }
    ptx::mbar_init(f_full, 1);
    ptx::mbar_init(f_empty, G::NCW);
    ptx::fence_barrier_init();
  }
  if (warp == G::NCW + 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t tile0 = tg, tstride = n_tg;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  double* part = p.partial + (size_t)blockIdx.x * (G::NB * G::MA + 3 * G::NB);  // [i_local][j] then Sd | Sq | L of the 64 own rows

  if (warp == G::NCW) {
    // =================================== producer ===================================
    if (lane == 0) {
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % G::YS);
        ptx::mbar_wait(&y_empty[st], (uint32_t)(((it / G::YS) & 1) ^ 1));  // first round: passes on the fresh barrier
        ptx::mbar_expect_tx(&y_full[st], (uint32_t)G::Y_STAGE_BYTES);
        const int t0 = (int)((tile0 + it * tstride) * G::KT);
        ptx::tma_load_2d(ysm + (size_t)st * G::Y_STAGE_BYTES, &tmap, t0, 0, &y_full[st]);
        ptx::tma_load_2d(ysm + (size_t)st * G::Y_STAGE_BYTES + G::Y_STAGE_BYTES / 2, &tmap, t0 + 16, 0, &y_full[st]);
      }
    }
  } else if (warp == G::NCW + 1) {
    // =================================== MMA issue ===================================
    if (lane == 0) {
      int since_flush = 0;
      uint32_t flushes = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int sl = (int)(it % G::SLOTS);
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 0] = clock64();
        ptx::mbar_wait(&o_full[sl], (uint32_t)((it / G::SLOTS) & 1));
        tc_fence_after();
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 1] = clock64();
        const uint32_t a0 = smem_u32(osm + (size_t)sl * G::SLOT_BYTES);
        const uint32_t b0 = a0 + S * G::A_DIGIT_BYTES;
        // Y digit pa against psi digits qb .. qb + m - 1 in ONE MMA of N = 64 m columns (the psi digits are adjacent 64-row blocks
        // of the slot, the level accumulators adjacent 64-column blocks): 8 MMAs per tile instead of 21, N up to 256.  Digit 0 of Y
        // touches every level: after a flush its MMAs re-initialise all accumulators.
#pragma unroll
        for (int pa = 0; pa < S; ++pa) {
#pragma unroll
          for (int qb = 0; qb < S - pa; qb += 4) {
            const int m = (S - pa - qb) < 4 ? (S - pa - qb) : 4;
            umma_i8_ss(tmem + (uint32_t)((pa + qb) * G::NB), make_desc_k32<LAYOUT>(a0 + pa * G::A_DIGIT_BYTES),
                       make_desc_k32<LAYOUT>(b0 + qb * G::B_DIGIT_BYTES), make_idesc(m * G::NB), (since_flush > 0 || pa > 0) ? 1u : 0u);
          }
        }
        umma_commit(&o_empty[sl]);
        if (TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS) trace[it * 8 + 2] = clock64();
        ++since_flush;
        if (since_flush == G::FLUSH_TILES || it + 1 == my_tiles) {
          umma_commit(f_full);
          if (it + 1 < my_tiles) {
            ptx::mbar_wait(f_empty, flushes & 1);
            tc_fence_after();
          }
          ++flushes;
          since_flush = 0;
        }
      }
    }
  } else {
    // =================================== converters ===================================
    // thread = (row r, sample group g of 8): warp = (row block rb of 16 rows, sample half sh); lane = (row offset, g & 1)
    const int rb = warp & 7, sh = warp >> 3;
    const int r = 16 * rb + (lane >> 1), g = 2 * sh + (lane & 1);
    const int ey = (r < p.n) ? p.rowexp[r] : 0;
    const double sc_y = scalbn(1.0, FRAC_BITS - ey), sc_psi = scalbn(1.0, FRAC_BITS - p.psi_exp);
    // Y stage: box g >> 1, row r, 16-byte chunks 4 (g & 1) + i at position chunk ^ (r & 7)
    const uint32_t y_off = (uint32_t)((g >> 1) * (G::Y_STAGE_BYTES / 2) + r * 128);
    // operand slot: this thread's 8 bytes of row r = half (g & 1) of the 16-byte chunk g >> 1
    const uint32_t a_off = slot_row_offset<LAYOUT>(r, g >> 1) + 8 * (g & 1);
    // psi: 4 samples 8 g + 4 q .. of the own row 64 half + rl, q = r >> 6 (warp-uniform): every thread does the same amount of work
    const int rl = r & 63, q = r >> 6, rp = 64 * half + rl;
    const bool from_regs = (q == half);  // the psi row is this thread's own Y row: the values are already in registers
    const uint32_t yp_off = (uint32_t)((g >> 1) * (G::Y_STAGE_BYTES / 2) + rp * 128);
    const uint32_t b_off = (uint32_t)(S * G::A_DIGIT_BYTES) + slot_row_offset<LAYOUT>(rl, g >> 1) + 8 * (g & 1) + 4 * q;
    double sd = 0.0;  // tanh: sum of tanh^2 ; other densities: sum of psi'
    // flush ownership: TMEM lane quarter warp & 3, 16 accumulator columns (warp >> 2) * 16 ..
    const int q4 = warp & 3, cg = warp >> 2;
    const int jrow = 32 * q4 + lane;
    const int ej = (jrow < p.n) ? p.rowexp[jrow] : 0;
#pragma unroll 1
    for (int c = 0; c < 16; ++c) part[(size_t)(cg * 16 + c) * G::MA + jrow] = 0.0;
    uint32_t flushes = 0;
    int since_flush = 0;

    auto flush = [&]() {
      ptx::mbar_wait(f_full, flushes & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(cg * 16);
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {
        int32_t c[S][8];
#pragma unroll
        for (int d = 0; d < S; ++d) tmem_ld8(taddr + (uint32_t)(d * G::NB + 8 * hc), c[d]);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const double v = scale_pow2(combine_levels(c[0][e], c[1][e], c[2][e], c[3][e], c[4][e], c[5][e]), ej + p.psi_exp + COMBINE_EXP);
          double* dst = part + (size_t)(cg * 16 + 8 * hc + e) * G::MA + jrow;
          *dst += v;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(f_empty);
      ++flushes;
    };

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int st = (int)(it % G::YS), sl = (int)(it % G::SLOTS);
      const int64_t t0 = (tile0 + it * tstride) * G::KT + 8 * g + 4 * q;  // first of this thread's 4 psi samples
      const bool tr = TRACE && blockIdx.x == 0 && it < I8_TRACE_SLOTS && warp == 0 && lane == 0;
      if (tr) trace[it * 8 + 4] = clock64();
      ptx::mbar_wait(&y_full[st], (uint32_t)((it / G::YS) & 1));
      if (tr) trace[it * 8 + 5] = clock64();
      double y[8], yp[4];
      {
        const unsigned char* yb = ysm + (size_t)st * G::Y_STAGE_BYTES + y_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(yb + ((((4 * (g & 1) + i) ^ (r & 7))) << 4));
          y[2 * i] = v.x; y[2 * i + 1] = v.y;
        }
        if (from_regs) {
#pragma unroll
          for (int i = 0; i < 4; ++i) yp[i] = q ? y[4 + i] : y[i];  // q is warp-uniform: a register select, no divergence
        } else {
          const unsigned char* ypb = ysm + (size_t)st * G::Y_STAGE_BYTES + yp_off;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(ypb + ((((4 * (g & 1) + 2 * q + i) ^ (rp & 7))) << 4));
            yp[2 * i] = v.x; yp[2 * i + 1] = v.y;
          }
        }
      }
      uint32_t wy[S][2], wp[S];
      digits8(y, sc_y, wy);
      // the slot is free once the MMAs of tile it - SLOTS have completed; the Y digits go out BEFORE the psi arithmetic, so that
      // most of this tile's shared-memory writes have completed long before the proxy fence below has to wait for them
      ptx::mbar_wait(&o_empty[sl], (uint32_t)(((it / G::SLOTS) & 1) ^ 1));
      if (tr) trace[it * 8 + 6] = clock64();
      unsigned char* ob = osm + (size_t)sl * G::SLOT_BYTES;
#pragma unroll
      for (int pb = 0; pb < S; ++pb)  // byte pb of the fixed-point integer = digit S - 1 - pb
        *reinterpret_cast<uint2*>(ob + (size_t)(S - 1 - pb) * G::A_DIGIT_BYTES + a_off) = make_uint2(wy[pb][0], wy[pb][1]);
      {
        double psi[4];
        if (NO_PSI) {
#pragma unroll
          for (int e = 0; e < 4; ++e) psi[e] = 0.5 * yp[e];
        } else if (DENS == DENS_TANH) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            psi[e] = tanh_psi(yp[e], p.dp, tab);
            sd = fma(psi[e], psi[e], sd);  // padding samples have y = 0, tanh = 0: no mask needed
          }
        } else {
          const bool partial_tile = (t0 + 4 > p.t_local);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            double fd = 0.0, dsd = 0.0, dsl = 0.0;
            density_eval<DENS, true, false, true>(yp[e], p.dp, tab, psi[e], fd, dsd, dsl);
            if (!partial_tile || t0 + e < p.t_local) sd += dsd;  // psi'(0) != 0: padding columns must not reach Sd
          }
        }
        digits4(psi, sc_psi, wp);
      }
#pragma unroll
      for (int pb = 0; pb < S; ++pb) *reinterpret_cast<uint32_t*>(ob + (size_t)(S - 1 - pb) * G::B_DIGIT_BYTES + b_off) = wp[pb];
      ptx::fence_proxy_async();  // generic-proxy stores before the tensor core's async-proxy reads
      __syncwarp();
      // The stage goes back to the producer only HERE, after the digit stores that consume every value read from it.  Released right
      // after the shared-memory loads were issued (round-2 first version), the arrival could overtake loads still in flight on the
      // sub-partition that also hosts the polling producer warp, and the refill replaced a few rows under them: a tile's worth of
      // error in G and Sd on about every second launch at T = 1e7 (profiles/lab: i8_lab mode 2 found it, run-to-run bit identity).
      if (lane == 0) { mbar_arrive(&y_empty[st]); mbar_arrive(&o_full[sl]); }
      if (tr) trace[it * 8 + 7] = clock64();
      ++since_flush;
      if (since_flush == G::FLUSH_TILES || it + 1 == my_tiles) { flush(); since_flush = 0; }
    }
    // Sd of the own rows: a row's 32 samples per tile live in lanes (g & 1) of the warps (sh, q) in {0, 1}^2
    sd += __shfl_xor_sync(0xffffffffu, sd, 1);
    if ((lane & 1) == 0) sdsm[(2 * sh + q) * 64 + rl] = sd;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * G::NCW) : "memory");
    if (tid < G::NB) {
      double v = (sdsm[tid] + sdsm[64 + tid]) + (sdsm[128 + tid] + sdsm[192 + tid]);
      if (DENS == DENS_TANH && !NO_PSI) {  // sum psi' = alpha (valid samples of this CTA - sum tanh^2)
        const int64_t last = tile0 + (my_tiles - 1) * tstride;
        const int64_t valid = my_tiles * G::KT - ((my_tiles > 0 && last == p.n_tiles - 1) ? (p.n_tiles * G::KT - p.t_local) : 0);
        v = p.dp.alpha * ((double)valid - v);
      }
      double* rs = part + G::NB * G::MA;
      rs[tid] = v;
      rs[G::NB + tid] = 0.0;
      rs[2 * G::NB + tid] = 0.0;
    }
  }
  __threadfence();  // this CTA's partial is visible device-wide before the CTA is counted
  tc_fence_before();
  __syncthreads();
  if (warp == G::NCW + 1) tmem_dealloc512(tmem);
  if (p.counter != nullptr) {
    if (tid == 0) {
      atomicAdd(p.counter, 1u);
      while ((int)(*reinterpret_cast<volatile unsigned int*>(p.counter) - p.target) < 0) __nanosleep(64);
      __threadfence();
    }
    __syncthreads();
    const int n = p.n, nn = n * n, total = nn + n;
    const int chunk = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * chunk, hi = lo + chunk < total ? lo + chunk : total;
    constexpr int PSZ = G::NB * G::MA + 3 * G::NB;
    // element e of the payload lives at mom[e]: Gr (n x n) is followed by Sd (n) in the moment buffer
    for (int e = lo + tid; e < hi; e += G::NTHREADS) {
      const int row = e < nn ? e / n : e - nn;                       // row of Gr / entry of Sd
      const int src = e < nn ? (row & 63) * G::MA + (e - row * n) : G::NB * G::MA + (row & 63);
      const double* pp = p.partial + (size_t)(row >> 6) * PSZ + src; // CTA 2 tg + half
      double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      int k = 0;
      for (; k + 8 <= n_tg; k += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += __ldcg(pp + (size_t)(k + u) * 2 * PSZ);
      }
      for (int u = 0; k < n_tg; ++k, ++u) acc[u] += __ldcg(pp + (size_t)k * 2 * PSZ);
      const double v = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
      if (!p.exchange) p.mom[e] = v;
      else
        for (int q = 0; q < p.px.nranks; ++q) p.px.peers.box[q][p2p_slot(p.px.parity, p.px.rank) + e] = v;
    }
    if (p.exchange) {
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        const unsigned int ticket = atomicAdd(p.counter2, 1u);
        if (ticket + 1u == p.target2) {  // the last CTA of this rank to finish pushing: the rank's contribution is complete everywhere
          __threadfence_system();
          for (int q = 0; q < p.px.nranks; ++q)
            atomicAdd_system(p.px.peers.flags[q] + p.px.parity * P2P_MAX_RANKS + p.px.rank, (unsigned int)P2P_PARTS);
        }
      }
      if (tid < p.px.nranks) {
        volatile unsigned int* f = p.px.peers.flags[p.px.rank] + p.px.parity * P2P_MAX_RANKS + tid;
        while ((int)(*f - p.px.expect) < 0) __nanosleep(32);
        __threadfence_system();
      }
      __syncthreads();
      const double* box = p.px.peers.box[p.px.rank] + p2p_slot(p.px.parity, 0);
      for (int e = lo + tid; e < hi; e += G::NTHREADS) {
        double acc = __ldcv(box + e);
        for (int q = 1; q < p.px.nranks; ++q) acc += __ldcv(box + (size_t)q * P2P_MAX_DOUBLES + e);
        p.mom[e] = acc;
      }
    }
  }
}

// e_j = bound_exponent(|w_j|_2 * max_t |x_t|_2): one warp per row
__global__ void row_exponent_kernel(const double* __restrict__ w, int n, const double* __restrict__ xstats, int* __restrict__ rowexp) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) { const double v = w[(size_t)row * n + k]; s = fma(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double xmax = sqrt(__longlong_as_double((long long)reinterpret_cast<const unsigned long long*>(xstats)[1]));
    rowexp[row] = bound_exponent(sqrt(s) * xmax * 1.0000001);
  }
}

}  // namespace i8
}  // namespace picard
