// =====================================================================================================
// rowblock.cuh -- the two kernels the core loop actually runs (core_solver.cu: every line-search try is a
// LOSS pass that keeps Y' in HBM; an accepted try is followed by the stored-Y gradient pass):
//
//   rb_loss_kernel<KP, DENS, MODE, WANT_SQ>   Y' = W X on DMMA, log-likelihood / y^2 row sums in registers, Y' stored
//       (LOSS: core.rs:124-127 with compute_loss core.rs:39-85) or just Y' (APPLY: solver.rs:140, whitening.rs:110,
//       solver.rs:199-214).  The A fragments (this warp's rows of W) live in REGISTERS for the whole kernel, so the
//       only shared-memory traffic of the contraction is one LDS.128 of the TMA-staged X tile per k-step.
//   rb_grady_kernel<NP, DENS, WANT_H, HAS_BIAS>   Gr = psi(Y) Y^T, Sd [, Hr = psi'(Y) (Y^2)^T, Sq] from the stored Y
//       (core.rs:215-221,264,274): the TMA tile of Y in shared memory is both the source of each warp's own elements
//       and the B operand of the DMMA contraction -- no staging copy, no CTA barrier.
//
// Both are ROW-BLOCK partitioned: a CTA owns RP rows of the output (of Y', resp. of Gr / Hr) and all columns, so the
// register-resident accumulators never exceed the register file whatever N is (N <= 256 here):
//   loss : KP = 64 -> RP = 64 ; KP = 128 -> RP = 128 (1 row block) ; KP = 256 -> RP = 64 (4 row blocks)   [A fragments: <= 64 doubles / thread]
//   grady: NP<=64 -> RP = NP ; 128 -> 128 (no H) / 64 (H) ; 256 -> 64 (no H) / 32 (H)   [accumulators: <= 64 doubles / thread]
// Row blocks of one sample tile run on different SMs and read the same X / Y tile (L2 absorbs the re-reads); no data
// is exchanged between CTAs.  Warps of a CTA are decoupled: stages are recycled by the last warp to release them.
//
// Sample permutation inside a 16-sample tile (free: every output is a sum over samples, or is stored back through the
// same map): with SWIZZLE_128B, 16-byte chunk q of row r sits at chunk position q ^ (r & 7).
//   loss : n-block nb, fragment column n  <->  sample 2 n + nb   (B fragment of k-step s = one LDS.128 of row k(s, j))
//   grady: lane j, half nb', slot pp      <->  sample 2 (2 j + nb') + pp
// both choices make every shared-memory access bank-conflict free.
// =====================================================================================================
#pragma once
#include <type_traits>

#include "pass.cuh"

namespace picard {

__host__ __device__ inline int rb_partial_size(int rp, int np, bool want_g, bool want_h) {
  return (want_g ? rp * np : 0) + (want_h ? rp * np : 0) + 3 * rp;
}

// Sum the per-CTA partials (CTA b = tile group b / nrb, row block b % nrb) in a fixed order into the compact moment
// buffer (leading dimension n).  Sections a mode does not produce are left untouched.
static __global__ void reduce_rb_kernel(const double* __restrict__ partial, int n_tg, int nrb, int rp, int np, int n, int want_g,
                                        int want_h, int want_l, double* __restrict__ mom) {
  const int psz = rb_partial_size(rp, np, want_g, want_h);
  const int64_t nn = (int64_t)n * n;
  const int64_t total = (want_g ? nn : 0) + (want_h ? nn : 0) + 3 * (int64_t)n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e;
    int src, row;
    int64_t dst;
    bool skip = false;
    if (want_g && r < nn) { row = (int)(r / n); src = (row % rp) * np + (int)(r % n); dst = mom_off_gr(n) + r; }
    else {
      if (want_g) r -= nn;
      if (want_h && r < nn) { row = (int)(r / n); src = rp * np + (row % rp) * np + (int)(r % n); dst = mom_off_hr(n) + r; }
      else {
        if (want_h) r -= nn;
        const int base = (want_g ? rp * np : 0) + (want_h ? rp * np : 0);
        const int sec = (int)(r / n);
        row = (int)(r % n);
        src = base + sec * rp + (row % rp);
        dst = (sec == 0 ? mom_off_sd(n) : (sec == 1 ? mom_off_sq(n) : mom_off_ll(n))) + row;
        if (sec == 0 && !want_g) skip = true;
        if (sec == 2 && !want_l) skip = true;
      }
    }
    if (skip) continue;
    const int rb = row / rp;
    // eight independent chains (the loads of a chain step are in flight together: the kernel is latency-bound), combined in a
    // fixed order: the result does not depend on the launch geometry
    const double* pp = partial + (size_t)rb * psz + src;
    const size_t kstride = (size_t)nrb * psz;
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int k = 0;
    for (; k + 8 <= n_tg; k += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += pp[(size_t)(k + u) * kstride];
    }
    for (int u = 0; k < n_tg; ++k, ++u) s[u] += pp[(size_t)k * kstride];
    mom[dst] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
}

// -----------------------------------------------------------------------------------------------------
// LOSS / APPLY
// -----------------------------------------------------------------------------------------------------
template <int KP>
struct RbLossGeom {
  static_assert(KP == 64 || KP == 128 || KP == 256, "register-resident A fragments are sized for KP = 64, 128 or 256");
  static constexpr int NWARPS = 8;
  static constexpr int MB = KP == 128 ? 2 : 1;   // 8-row blocks per warp: MB * KP / 4 = 64 A-fragment doubles per thread
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int RP = 8 * MB * NWARPS;     // rows of Y' per CTA
  static constexpr int KS = KP / 4;              // k-steps
  static constexpr int BT = 16;
  static constexpr int STAGES = KP == 256 ? 4 : 6;
  static constexpr int MIN_BLOCKS = KP == 64 ? 2 : 1;  // KP = 64: 16 A-fragment doubles per thread, two CTAs per SM
  static constexpr bool BIG_TAB = KP >= 128;           // 80 KB of density tables next to the stages (one CTA per SM anyway)
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * KP * BT * 8 + (size_t)dmath::Tab<BIG_TAB>::DOUBLES * 8 + 128;
};

template <int KP, int DENS, int MODE, bool WANT_SQ>
__global__ void __launch_bounds__(RbLossGeom<KP>::NTHREADS, RbLossGeom<KP>::MIN_BLOCKS)
rb_loss_kernel(const __grid_constant__ CUtensorMap tmap, const PassParams p, const int nrb) {
  using G = RbLossGeom<KP>;
  constexpr bool APPLY = (MODE == PASS_APPLY);
  constexpr bool NEED_TAB = !APPLY && (DENS == DENS_TANH || DENS == DENS_EXP);
  constexpr int MB = G::MB, KS = G::KS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  constexpr bool BIG = G::BIG_TAB;
  double* tab = xs + G::STAGES * KP * G::BT;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tab + dmath::Tab<BIG>::DOUBLES);
  int* cnt = reinterpret_cast<int*>(bar + G::STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = lane & 3, c = lane >> 2;
  const int rb = blockIdx.x % nrb, tg = blockIdx.x / nrb, n_tg = gridDim.x / nrb;
  const int r0 = rb * G::RP;

  if (NEED_TAB) load_density_tables<BIG>(tab, DENS == DENS_TANH, tid, G::NTHREADS);
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int s = 0; s < G::STAGES; ++s) { ptx::mbar_init(&bar[s], 1); cnt[s] = 0; }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  const int64_t tile0 = tg, tstride = n_tg;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  constexpr uint32_t STAGE_BYTES = KP * G::BT * 8;
  if (tid == 0) {
    for (int s = 0; s < G::STAGES && s < my_tiles; ++s) {
      ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
      ptx::tma_load_2d(xs + s * KP * G::BT, &tmap, (int)((tile0 + s * tstride) * G::BT), 0, &bar[s]);
    }
  }

  // ---- A fragments: this warp's rows of W, all KP columns, in registers.  k-step s, lane j <-> k = 8 (s / 2) + 2 j + (s & 1)
  double areg[MB][KS];
  double brow[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    const int row = r0 + 8 * (MB * warp + mb) + c;
    brow[mb] = (p.bias != nullptr && row < p.n_out) ? p.bias[row] : 0.0;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int k = 8 * (s >> 1) + 2 * j + (s & 1);
      areg[mb][s] = (row < p.n_out && k < p.n_in) ? p.w[(size_t)row * p.ldw + k] : 0.0;
    }
  }
  double sq[MB], sl[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) sq[mb] = sl[mb] = 0.0;
  // B fragment of k-step s: row k(s, j), 16-byte chunk c ^ (k & 7) = c ^ (2 j + (s & 1))  ->  samples 2c (nb 0), 2c + 1 (nb 1)
  int xoff[2];
#pragma unroll
  for (int b = 0; b < 2; ++b) xoff[b] = (2 * j + b) * G::BT + ((c ^ (2 * j + b)) << 1);  // + (s / 2) * 8 * BT

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = (int)(it % G::STAGES);
    const uint32_t parity = (uint32_t)((it / G::STAGES) & 1);
    const int64_t t0 = (tile0 + it * tstride) * G::BT;
    const bool partial_tile = (t0 + G::BT > p.t_local);
    const double* xst = xs + stage * KP * G::BT;
    ptx::mbar_wait(&bar[stage], parity);

    double acc[MB][2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const double2 b = *reinterpret_cast<const double2*>(xst + (s >> 1) * 8 * G::BT + xoff[s & 1]);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        ptx::dmma(acc[mb][0][0], acc[mb][0][1], areg[mb][s], b.x);
        ptx::dmma(acc[mb][1][0], acc[mb][1][1], areg[mb][s], b.y);
      }
    }
    // the stage is free as soon as the contraction has consumed it
    ptx::stage_release<G::NWARPS>(&cnt[stage], lane, [&] {
      if (it + G::STAGES < my_tiles) {
        ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
        ptx::tma_load_2d(xs + stage * KP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
      }
    });
    // acc[mb][nb][pp] <-> row 8 (MB warp + mb) + c, sample 2 (2 j + pp) + nb.
    // Two instantiations of the epilogue: interior tiles (every sample valid: no per-element bounds logic -- the epilogue is
    // issue-bound, and every instruction saved shortens the time both warps of a scheduler spend off the DMMA pipe) and the
    // one partial tile at the end of the shard.
    auto epilogue = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            double y = APPLY ? acc[mb][nb][pp] - brow[mb] : acc[mb][nb][pp];
            bool valid = true;
            if (!FULL) {
              const int64_t t = t0 + 4 * j + 2 * pp + nb;
              valid = t < p.t_local;
              if (!valid) y = 0.0;
            }
            acc[mb][nb][pp] = y;
            if (!APPLY) {
              double f = 0.0, fd = 0.0, dsd = 0.0, dsl = 0.0;
              if (FULL) {
                density_eval<DENS, false, true, BIG>(y, p.dp, tab, f, fd, dsd, sl[mb]);
              } else {
                density_eval<DENS, false, true, BIG>(y, p.dp, tab, f, fd, dsd, dsl);
                if (valid) sl[mb] += dsl;  // loglik(0) != 0: padding columns must not reach L
              }
              if (WANT_SQ) sq[mb] = fma(y, y, sq[mb]);
            }
          }
      if (p.out != nullptr) {  // Y' kept in HBM (LOSS: for the stored-Y gradient pass; APPLY: the result)
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const int row = r0 + 8 * (MB * warp + mb) + c;
          if (row < p.n_out) {
            double* dst = p.out + (size_t)row * p.ld_out + t0 + 4 * j;
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              if (FULL) {
                *reinterpret_cast<double2*>(dst + 2 * pp) = make_double2(acc[mb][0][pp], acc[mb][1][pp]);
              } else {
                const int64_t t = t0 + 4 * j + 2 * pp;
                if (t + 1 < p.t_local) *reinterpret_cast<double2*>(dst + 2 * pp) = make_double2(acc[mb][0][pp], acc[mb][1][pp]);
                else if (t < p.t_local) dst[2 * pp] = acc[mb][0][pp];
              }
            }
          }
        }
      }
    };
    if (!partial_tile) epilogue(std::true_type{});
    else epilogue(std::false_type{});
  }

  if (!APPLY) {
    double* rs = p.partial + (size_t)blockIdx.x * rb_partial_size(G::RP, KP, false, false);
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      double b = sq[mb], l = sl[mb];
      b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
      l += __shfl_xor_sync(0xffffffffu, l, 1); l += __shfl_xor_sync(0xffffffffu, l, 2);
      if (j == 0) {
        const int row = 8 * (MB * warp + mb) + c;
        rs[row] = 0.0; rs[G::RP + row] = b; rs[2 * G::RP + row] = l;
      }
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// stored-Y gradient moments
// -----------------------------------------------------------------------------------------------------
template <int NP, bool WANT_H>
struct RbGradYGeom {
  static constexpr int NWARPS = NP >= 64 ? 8 : NP / 8;
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int CW = (NP == 256 && WANT_H) ? 2 : 1;          // warps sharing a row group, splitting the columns
  static constexpr int MB = (NP == 128 && !WANT_H) ? 2 : 1;         // 8-row blocks per warp
  static constexpr int RP = 8 * MB * NWARPS / CW;                   // rows of Gr / Hr per CTA
  static constexpr int NRB = NP / RP;
  static constexpr int NBW = NP / CW / 8;                           // 8-column blocks per warp
  static constexpr int BT = 16;
  static constexpr int STAGES = NP == 256 ? 4 : 6;
  static constexpr int MIN_BLOCKS = NP >= 128 ? 1 : (NP == 64 ? 2 : (NP == 32 ? 4 : 8));
  static constexpr bool BIG_TAB = NP >= 128;  // only the exp part (16 KB) is staged: no log-likelihood in this kernel
  static constexpr size_t TAB_BYTES = (size_t)(BIG_TAB ? dmath::Tab<true>::EXP_N : dmath::TAB_DOUBLES) * 8;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * NP * BT * 8 + TAB_BYTES + (size_t)NP * 8 + 128;
  static_assert(MB * NBW * (WANT_H ? 2 : 1) <= 32, "accumulators exceed 128 registers per thread");
};

template <int NP, int DENS, bool WANT_H, bool HAS_BIAS>
__global__ void __launch_bounds__(RbGradYGeom<NP, WANT_H>::NTHREADS, RbGradYGeom<NP, WANT_H>::MIN_BLOCKS)
rb_grady_kernel(const __grid_constant__ CUtensorMap tmap, const PassParams p, const int nrb) {
  using G = RbGradYGeom<NP, WANT_H>;
  constexpr int MB = G::MB, NBW = G::NBW, CW = G::CW;
  constexpr bool NEED_TAB = (DENS == DENS_TANH || DENS == DENS_EXP);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* ysm = reinterpret_cast<double*>(smem_raw);
  constexpr bool BIG = G::BIG_TAB;
  double* tab = ysm + G::STAGES * NP * G::BT;
  double* bs = tab + G::TAB_BYTES / 8;
  uint64_t* bar = reinterpret_cast<uint64_t*>(bs + NP);
  int* cnt = reinterpret_cast<int*>(bar + G::STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = lane & 3, c = lane >> 2;
  const int rb = blockIdx.x % nrb, tg = blockIdx.x / nrb, n_tg = gridDim.x / nrb;
  const int rg = warp / CW, ch = warp % CW;
  const int row0 = rb * G::RP + 8 * MB * rg;  // first row of this warp (multiple of 8)
  const int col0 = ch * NBW * 8;              // first column of this warp

  if (NEED_TAB) load_density_tables<BIG>(tab, false, tid, G::NTHREADS);
  if (HAS_BIAS)
    for (int i = tid; i < NP; i += G::NTHREADS) bs[i] = (p.bias != nullptr && i < p.n_out) ? p.bias[i] : 0.0;
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int s = 0; s < G::STAGES; ++s) { ptx::mbar_init(&bar[s], 1); cnt[s] = 0; }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  const int64_t tile0 = tg, tstride = n_tg;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  constexpr uint32_t STAGE_BYTES = NP * G::BT * 8;
  if (tid == 0) {
    for (int s = 0; s < G::STAGES && s < my_tiles; ++s) {
      ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
      ptx::tma_load_2d(ysm + s * NP * G::BT, &tmap, (int)((tile0 + s * tstride) * G::BT), 0, &bar[s]);
    }
  }

  double gacc[MB][NBW][2];
  double hacc[WANT_H ? MB : 1][WANT_H ? NBW : 1][2];
#pragma unroll
  for (int a = 0; a < MB; ++a)
#pragma unroll
    for (int b = 0; b < NBW; ++b) gacc[a][b][0] = gacc[a][b][1] = 0.0;
#pragma unroll
  for (int a = 0; a < (WANT_H ? MB : 1); ++a)
#pragma unroll
    for (int b = 0; b < (WANT_H ? NBW : 1); ++b) hacc[a][b][0] = hacc[a][b][1] = 0.0;
  double sd[MB], sq[MB], brow[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) { sd[mb] = sq[mb] = 0.0; brow[mb] = HAS_BIAS ? bs[row0 + 8 * mb + c] : 0.0; }

  // lane-constant offsets (doubles); every row index used below is congruent to c modulo 8
  int eoff[2], boff[2];
#pragma unroll
  for (int nbp = 0; nbp < 2; ++nbp) {
    eoff[nbp] = (row0 + c) * G::BT + (((2 * j + nbp) ^ c) << 1);  // + mb * 8 * BT   : own elements
    boff[nbp] = (col0 + c) * G::BT + (((2 * j + nbp) ^ c) << 1);  // + nbg * 8 * BT  : B fragments
  }

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = (int)(it % G::STAGES);
    const uint32_t parity = (uint32_t)((it / G::STAGES) & 1);
    const int64_t t0 = (tile0 + it * tstride) * G::BT;
    const bool partial_tile = (t0 + G::BT > p.t_local);
    const double* yt = ysm + stage * NP * G::BT;
    ptx::mbar_wait(&bar[stage], parity);

    double psi[MB][2][2], psd[WANT_H ? MB : 1][2][2];
    // interior tiles (every sample valid) take the instantiation without per-element bounds logic
    auto own_elements = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nbp = 0; nbp < 2; ++nbp) {
          const double2 v = *reinterpret_cast<const double2*>(yt + eoff[nbp] + mb * 8 * G::BT);
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            double y = pp ? v.y : v.x;  // out-of-range columns / rows are zero-filled by the TMA unit
            bool valid = true;
            if (!FULL) valid = (t0 + 2 * (2 * j + nbp) + pp) < p.t_local;
            if (HAS_BIAS) y = valid ? y - brow[mb] : 0.0;
            double f = 0.0, fd = 0.0, dsd = 0.0, dsl = 0.0;
            if (FULL) {
              density_eval<DENS, true, false, BIG>(y, p.dp, tab, f, fd, sd[mb], dsl);
            } else {
              density_eval<DENS, true, false, BIG>(y, p.dp, tab, f, fd, dsd, dsl);
              if (valid) sd[mb] += dsd;  // psi'(0) != 0: padding columns must not reach Sd
            }
            psi[mb][nbp][pp] = f;
            if (WANT_H) { psd[mb][nbp][pp] = fd; sq[mb] = fma(y, y, sq[mb]); }
          }
        }
    };
    // measured: a win at NP = 256 (8.44 -> 8.13 ms), a loss at NP = 128 (11.39 -> 11.60 ms: the second copy of the unrolled
    // density code costs more than the bounds logic it removes), so the specialisation is compiled for NP != 128 only
    if (NP != 128 && !partial_tile) own_elements(std::true_type{});
    else own_elements(std::false_type{});
#pragma unroll
    for (int nbp = 0; nbp < 2; ++nbp)
#pragma unroll
      for (int nbg = 0; nbg < NBW; ++nbg) {
        double2 b = *reinterpret_cast<const double2*>(yt + boff[nbp] + nbg * 8 * G::BT);
        if (HAS_BIAS) {  // centring folded in (covariance pass of the whitening step only)
          const double bc = bs[col0 + 8 * nbg + c];
          const int64_t t = t0 + 2 * (2 * j + nbp);
          b.x = (!partial_tile || t < p.t_local) ? b.x - bc : 0.0;
          b.y = (!partial_tile || t + 1 < p.t_local) ? b.y - bc : 0.0;
        }
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          ptx::dmma(gacc[mb][nbg][0], gacc[mb][nbg][1], psi[mb][nbp][0], b.x);
          ptx::dmma(gacc[mb][nbg][0], gacc[mb][nbg][1], psi[mb][nbp][1], b.y);
        }
        if (WANT_H) {
          const double bx2 = b.x * b.x, by2 = b.y * b.y;
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            ptx::dmma(hacc[mb][nbg][0], hacc[mb][nbg][1], psd[mb][nbp][0], bx2);
            ptx::dmma(hacc[mb][nbg][0], hacc[mb][nbg][1], psd[mb][nbp][1], by2);
          }
        }
      }
    ptx::stage_release<G::NWARPS>(&cnt[stage], lane, [&] {
      if (it + G::STAGES < my_tiles) {
        ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
        ptx::tma_load_2d(ysm + stage * NP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
      }
    });
  }

  double* part = p.partial + (size_t)blockIdx.x * rb_partial_size(G::RP, NP, true, WANT_H);
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int nbg = 0; nbg < NBW; ++nbg) {
      const int lrow = 8 * (MB * rg + mb) + c, col = col0 + 8 * nbg + 2 * j;
      *reinterpret_cast<double2*>(part + lrow * NP + col) = make_double2(gacc[mb][nbg][0], gacc[mb][nbg][1]);
      if (WANT_H) *reinterpret_cast<double2*>(part + G::RP * NP + lrow * NP + col) = make_double2(hacc[mb][nbg][0], hacc[mb][nbg][1]);
    }
  double* rs = part + G::RP * NP + (WANT_H ? G::RP * NP : 0);
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    double a = sd[mb], b = sq[mb];
    a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
    if (j == 0 && ch == 0) {
      const int lrow = 8 * (MB * rg + mb) + c;
      rs[lrow] = a; rs[G::RP + lrow] = b; rs[2 * G::RP + lrow] = 0.0;
    }
  }
}

}  // namespace picard
