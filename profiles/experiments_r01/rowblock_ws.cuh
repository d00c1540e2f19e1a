// =====================================================================================================
// rowblock_ws.cuh -- warp-specialised form of the LOSS / APPLY row-block kernel (rowblock.cuh) for KP >= 128.
// EXPERIMENT, not the default path (PICARD_RB_WS=1 selects it): parity-green, measured SLOWER than rb_loss_kernel.
//
// Background (phase-trace build, profiles/rb_trace.py, profiles/rb_trace_r01g_*.json).  DMMA and DFMA share one FP64 pipe per
// scheduler.  In rb_loss_kernel every warp alternates between a DMMA phase (128 DMMAs, ~4100 cycles when the two warps of a
// scheduler alternate on the pipe) and a density/store phase (~750 cycles).  The two warps of a scheduler run phase-locked
// whatever their initial offset (a head start of 64 DMMAs decays within ~16 tiles): their density phases overlap by ~60 %,
// and in that joint phase the pipe is about half idle (the phase is issue-bound: ~265 instructions per warp, 120 of them
// FP64).  Forcing the density phases to take turns is much worse (11.1 -> 16.3 ms): next to the other warp's DMMA phase a
// density phase gets about ONE FP64 instruction per 16-cycle DMMA and lasts 3450 cycles instead of 750.
//
// This kernel tries the other way out: split the roles.  8 DMMA warps (two per scheduler, always in their DMMA phase, 128
// registers of W fragments each) write Y' tiles into a shared-memory ring; 16 density warps (four per scheduler) evaluate the
// log-likelihood row sums from the ring and ONE thread sends each tile to HBM with a TMA store (which also clips the ragged
// last tile and the rows >= n_out).  The register file is re-partitioned with setmaxnreg inside the CTA allocation
// (768 threads x 80 at launch -> DMMA warps 160, density warps 40; the pool is the CTA's own: a sum above the launch
// allocation deadlocks in setmaxnreg.inc).
//
// Measured on B200, N = 128, T = 1e7, tanh (profiles/ws_ab.sh): 13.9 ms with 8 density warps, 12.5 ms with 16, against
// 10.85 ms for rb_loss_kernel; N = 256 exp: 8.5 vs 7.9 ms.  Density instructions interleaved one by one with DMMAs cost the
// pipe more than they occupy it, and a density warp that gets one pipe slot per round of the scheduler's warps becomes the
// critical path.  The joint, phase-locked density phase of rb_loss_kernel is the better schedule on this machine.
// =====================================================================================================
#pragma once
#include "rowblock.cuh"

namespace picard {

namespace ptx {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA: 2-D tiled tensor store shared -> global (SASS: UTMASTG), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
}  // namespace ptx

constexpr int WS_REG_DMMA = 160;  // setmaxnreg moves registers inside the CTA allocation (768 threads x 80 = 61440): 8 x 32 x 160 + 16 x 32 x 40 = 61440
constexpr int WS_REG_DENS = 40;

template <int KP>
struct RbLossWsGeom {
  static_assert(KP == 128 || KP == 256, "warp-specialised LOSS / APPLY kernel: KP = 128 or 256");
  static constexpr int NDW = 8;                  // DMMA warps
  static constexpr int NEW = 16;                 // density warps
  static constexpr int NTHREADS = 32 * (NDW + NEW);
  static constexpr int MB = KP == 128 ? 2 : 1;   // 8-row blocks per DMMA warp: MB * KP / 4 = 64 A-fragment doubles per thread
  static constexpr int RP = 8 * MB * NDW;        // rows of Y' per CTA
  static constexpr int KS = KP / 4;
  static constexpr int BT = 16;
  static constexpr int STAGES = KP == 256 ? 3 : 5;
  static constexpr int NYB = 2;                  // Y' tiles in flight between the two roles
  static constexpr int TPR = 32 * NEW / RP;      // density threads per row (4 or 8)
  static constexpr int CPT = 8 / TPR;            // 16-byte chunks (2 samples) per density thread and tile
  static constexpr bool BIG_TAB = true;
  static constexpr size_t XS_BYTES = (size_t)STAGES * KP * BT * 8;
  static constexpr size_t YS_BYTES = (size_t)NYB * RP * BT * 8;
  static constexpr size_t TAB_BYTES = (size_t)dmath::Tab<BIG_TAB>::DOUBLES * 8;
  static constexpr size_t SMEM_BYTES = XS_BYTES + YS_BYTES + TAB_BYTES + 256;
};

template <int KP, int DENS, int MODE, bool WANT_SQ>
__global__ void __launch_bounds__(RbLossWsGeom<KP>::NTHREADS, 1)
rb_loss_ws_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const PassParams p,
                  const int nrb) {
  using G = RbLossWsGeom<KP>;
  constexpr bool APPLY = (MODE == PASS_APPLY);
  constexpr bool NEED_TAB = !APPLY && (DENS == DENS_TANH || DENS == DENS_EXP);
  constexpr bool BIG = G::BIG_TAB;
  constexpr int MB = G::MB, KS = G::KS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* ys = xs + G::STAGES * KP * G::BT;          // 1024-byte aligned: every stage / tile is a multiple of 1 KB
  double* tab = ys + G::NYB * G::RP * G::BT;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tab + dmath::Tab<BIG>::DOUBLES);  // X ring: TMA transaction barriers
  uint64_t* yfull = bar + G::STAGES;                 // Y' ring: tile written (NDW arrivals)
  uint64_t* yempty = yfull + G::NYB;                 // Y' ring: tile consumed (density warps / the storing thread)
  int* cnt = reinterpret_cast<int*>(yempty + G::NYB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rb = blockIdx.x % nrb, tg = blockIdx.x / nrb, n_tg = gridDim.x / nrb;
  const int r0 = rb * G::RP;
  const bool store_y = p.out != nullptr;
  // APPLY: the density role is one thread that issues the TMA stores
  const int n_consumers = APPLY ? 1 : G::NEW;

  if (NEED_TAB) load_density_tables<BIG>(tab, DENS == DENS_TANH, tid, G::NTHREADS);
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    if (store_y) ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < G::STAGES; ++s) { ptx::mbar_init(&bar[s], 1); cnt[s] = 0; }
    for (int s = 0; s < G::NYB; ++s) { ptx::mbar_init(&yfull[s], G::NDW); ptx::mbar_init(&yempty[s], n_consumers); }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  const int64_t tile0 = tg, tstride = n_tg;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  constexpr uint32_t STAGE_BYTES = KP * G::BT * 8;

  if (warp >= G::NEW) {
    // =================================== DMMA warps ===================================
    ptx::setmaxnreg_inc<WS_REG_DMMA>();
    const int dw = warp - G::NEW;
    const int j = lane & 3, c = lane >> 2;
    if (dw == 0 && lane == 0) {
      for (int s = 0; s < G::STAGES && s < my_tiles; ++s) {
        ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
        ptx::tma_load_2d(xs + s * KP * G::BT, &tmap, (int)((tile0 + s * tstride) * G::BT), 0, &bar[s]);
      }
    }
    // A fragments: this warp's rows of W, all KP columns, in registers.  k-step s, lane j <-> k = 8 (s / 2) + 2 j + (s & 1)
    double areg[MB][KS];
    double brow[MB];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int row = r0 + 8 * (MB * dw + mb) + c;
      brow[mb] = (APPLY && p.bias != nullptr && row < p.n_out) ? p.bias[row] : 0.0;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int k = 8 * (s >> 1) + 2 * j + (s & 1);
        areg[mb][s] = (row < p.n_out && k < p.n_in) ? p.w[(size_t)row * p.ldw + k] : 0.0;
      }
    }
    // B fragment of k-step s: row k(s, j), 16-byte chunk c ^ (k & 7) = c ^ (2 j + (s & 1))  ->  samples 2c (nb 0), 2c + 1 (nb 1)
    int xoff[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) xoff[b] = (2 * j + b) * G::BT + ((c ^ (2 * j + b)) << 1);  // + (s / 2) * 8 * BT
    // Y' tile: acc[mb][nb][pp] <-> row 8 (MB dw + mb) + c, samples 4 j + 2 pp + nb  =  16-byte chunk 2 j + pp of that row,
    // stored at chunk position (2 j + pp) ^ (row & 7) (the SWIZZLE_128B pattern the TMA store expects): conflict-free STS.128
    int yoff[MB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) yoff[mb][pp] = (8 * (MB * dw + mb) + c) * G::BT + (((2 * j + pp) ^ c) << 1);

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int stage = (int)(it % G::STAGES);
      const uint32_t parity = (uint32_t)((it / G::STAGES) & 1);
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
      ptx::stage_release<G::NDW>(&cnt[stage], lane, [&] {
        if (it + G::STAGES < my_tiles) {
          ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
          ptx::tma_load_2d(xs + stage * KP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
        }
      });

      const int yb = (int)(it % G::NYB);
      ptx::mbar_wait(&yempty[yb], (uint32_t)(((it / G::NYB) & 1) ^ 1));  // first round: passes on the fresh barrier
      double* yt = ys + yb * G::RP * G::BT;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          double v0 = acc[mb][0][pp], v1 = acc[mb][1][pp];
          if (APPLY) { v0 -= brow[mb]; v1 -= brow[mb]; }
          *reinterpret_cast<double2*>(yt + yoff[mb][pp]) = make_double2(v0, v1);
        }
      ptx::fence_proxy_async();  // generic-proxy writes of the tile before the async-proxy (TMA store) read
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&yfull[yb]);
    }
  } else {
    // =================================== density warps ===================================
    ptx::setmaxnreg_dec<WS_REG_DENS>();
    if (APPLY && warp != 0) return;
    const int et = tid;                       // 0 .. 32 NEW - 1
    const int row = et / G::TPR, sub = et % G::TPR;
    double sq = 0.0, sl = 0.0;
    int coff[G::CPT];
#pragma unroll
    for (int k = 0; k < G::CPT; ++k) coff[k] = row * G::BT + (((sub * G::CPT + k) ^ (row & 7)) << 1);

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int yb = (int)(it % G::NYB);
      const int64_t t0 = (tile0 + it * tstride) * G::BT;
      const double* yt = ys + yb * G::RP * G::BT;
      ptx::mbar_wait(&yfull[yb], (uint32_t)((it / G::NYB) & 1));
      if (store_y && tid == 0) {
        ptx::tma_store_2d(&tmap_out, yt, (int)t0, r0);
        ptx::bulk_commit();
      }
      if (!APPLY) {
        const bool partial_tile = (t0 + G::BT > p.t_local);
        if (!partial_tile) {
#pragma unroll
          for (int k = 0; k < G::CPT; ++k) {
            const double2 v = *reinterpret_cast<const double2*>(yt + coff[k]);
            double f = 0.0, fd = 0.0, dsd = 0.0;
            density_eval<DENS, false, true, BIG>(v.x, p.dp, tab, f, fd, dsd, sl);
            density_eval<DENS, false, true, BIG>(v.y, p.dp, tab, f, fd, dsd, sl);
            if (WANT_SQ) { sq = fma(v.x, v.x, sq); sq = fma(v.y, v.y, sq); }
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < G::CPT; ++k) {
            const double2 v = *reinterpret_cast<const double2*>(yt + row * G::BT + (((sub * G::CPT + k) ^ (row & 7)) << 1));
            const int64_t t = t0 + 2 * (sub * G::CPT + k);
            double f = 0.0, fd = 0.0, dsd = 0.0, d0 = 0.0, d1 = 0.0;
            density_eval<DENS, false, true, BIG>(v.x, p.dp, tab, f, fd, dsd, d0);
            density_eval<DENS, false, true, BIG>(v.y, p.dp, tab, f, fd, dsd, d1);
            if (t < p.t_local) { sl += d0; if (WANT_SQ) sq = fma(v.x, v.x, sq); }  // loglik(0) != 0: padding must not reach L
            if (t + 1 < p.t_local) { sl += d1; if (WANT_SQ) sq = fma(v.y, v.y, sq); }
          }
        }
      }
      if (store_y && tid == 0) ptx::bulk_wait_read0();  // the TMA unit has read the tile: it may be overwritten
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&yempty[yb]);
    }
    if (store_y && tid == 0) ptx::bulk_wait0();         // all stores complete before the CTA retires
    if (!APPLY) {
      double* rs = p.partial + (size_t)blockIdx.x * rb_partial_size(G::RP, KP, false, false);
#pragma unroll
      for (int o = 1; o < G::TPR; o <<= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
      }
      if (sub == 0) { rs[row] = 0.0; rs[G::RP + row] = sq; rs[2 * G::RP + row] = sl; }
    }
  }
}

}  // namespace picard
