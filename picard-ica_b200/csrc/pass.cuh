// =====================================================================================================
// pass.cuh -- the per-iteration pass over the N x T sample matrix as ONE fused sm_100a kernel
// (replaces core.rs:215-221,226,264,274 + density.rs + the `transform.dot(y)` / compute_loss pair of
// core.rs:124-127, and, in APPLY mode, solver.rs:140 / whitening.rs:110 / solver.rs:199-214).
//
// Per tile of BT = 16 samples, one CTA:
//   TMA   : X tile [n_in x 16] f64 -> shared memory (cp.async.bulk.tensor, SWIZZLE_128B, 3-stage mbarrier
//           ring, out-of-bounds rows/columns zero-filled by the TMA unit),
//   step 1: Y = W X - bias on the FP64 tensor path (DMMA m8n8k4), W resident in shared memory,
//   step 2: psi, psi', log-likelihood in registers on the accumulator fragments (density.cuh), row sums
//           Sd = sum psi', Sq = sum y^2, L = sum loglik kept per thread,
//   step 3: Gr += psi(Y) Y^T  [Hr += psi'(Y) (Y^2)^T]  again on DMMA: the A operand is the psi fragment
//           straight from registers (the contraction index t may be permuted freely, so the step-1
//           accumulator layout IS a valid A-fragment layout), the B operand is Y staged once through
//           shared memory; the N x N accumulators stay in registers for the whole kernel.
// Each CTA writes one partial of the packed moment buffer; reduce_partials() sums them in a fixed order
// (deterministic), after which a single allreduce combines GPUs.
//
// FP64 has no tcgen05 kind (tcgen05.mma is f16/tf32/f8/f6/f4/i8 only) and no TMEM accumulator format, so
// the tensor path for f64 on sm_100a is DMMA.8x8x4; measured 37.2 TFLOP/s vs 34.1 for DFMA, same pipe.
// =====================================================================================================
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "density.cuh"

namespace picard {

template <int NP>
struct PassGeom {
  static_assert(NP == 8 || NP == 16 || NP == 32 || NP == 64 || NP == 128, "unsupported padded size");
  static constexpr int NWARPS = NP >= 64 ? 8 : NP / 8;
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int MB = NP / (8 * NWARPS);  // 8-row blocks of Y / G owned by one warp
  static constexpr int NB = NP / 8;             // 8-column blocks of G
  static constexpr int BT = 16;                 // samples per tile = one 128-byte swizzle row
  static constexpr int STAGES = 3;
  static constexpr int WPITCH = NP + 4;         // +32 B: conflict-free A-fragment loads
  static constexpr int YPITCH = BT + 4;
  static constexpr int MIN_BLOCKS = NP >= 128 ? 1 : (NP == 64 ? 2 : (NP == 32 ? 4 : 8));
  static constexpr size_t XS_BYTES = (size_t)STAGES * NP * BT * 8;
  static constexpr size_t WS_BYTES = (size_t)NP * WPITCH * 8;
  static constexpr size_t YS_BYTES = (size_t)2 * NP * YPITCH * 8;
  static constexpr size_t TAB_BYTES = (size_t)dmath::TAB_DOUBLES * 8;  // exp / log lookup tables (density.cuh)
  static constexpr size_t SMEM_BYTES = XS_BYTES + WS_BYTES + YS_BYTES + (size_t)NP * 8 + TAB_BYTES + 64;
};

constexpr int PASS_MAX_BLOCKS_PER_SM = 8;  // grid <= sm_count * this: bounds the per-CTA partial workspace

// packed partial layout of one CTA (padded to NP): [G NP^2 if WANT_G][H NP^2 if WANT_H][SD NP][SQ NP][LL NP]
__host__ __device__ inline int pass_partial_size(int np, bool want_g, bool want_h) {
  return (want_g ? np * np : 0) + (want_h ? np * np : 0) + 3 * np;
}
// compact reduced layout (leading dimension n): [GR n^2][SD n][SQ n][LL n][HR n^2] -- ordered so that what a
// pass mode produces is one contiguous range: loss-only = [SQ, LL]; grad = [GR, SD, SQ] (+HR); fused = all.
__host__ __device__ inline int64_t mom_size(int n) { return 2 * (int64_t)n * n + 3 * (int64_t)n; }
__host__ __device__ inline int64_t mom_off_gr(int n) { return 0; }
__host__ __device__ inline int64_t mom_off_sd(int n) { return (int64_t)n * n; }
__host__ __device__ inline int64_t mom_off_sq(int n) { return (int64_t)n * n + n; }
__host__ __device__ inline int64_t mom_off_ll(int n) { return (int64_t)n * n + 2 * (int64_t)n; }
__host__ __device__ inline int64_t mom_off_hr(int n) { return (int64_t)n * n + 3 * (int64_t)n; }

struct PassParams {
  const double* w;     // (n_out x n_in), leading dimension ldw, device
  const double* bias;  // (n_out) or nullptr: y = W x - bias
  int n_out, n_in, ldw;
  int64_t t_local;     // samples in this shard
  int64_t n_tiles;     // ceil(t_local / 16)
  DensParams dp;
  double* partial;     // [gridDim.x][pass_partial_size]
  double* out;         // APPLY: (n_out x t_local), leading dimension ld_out (even)
  int64_t ld_out;
};

// lookup tables of density.cuh in global memory (one copy per translation unit); staged into shared memory per CTA
static __device__ const double g_exp_tab[dmath::EXP_TAB_N] = PICARD_EXP_TAB_INIT;
static __device__ const double g_log_tab[dmath::LOG_TAB_N] = PICARD_LOG_TAB_INIT;
static __device__ const double g_exp_tab_big[dmath::Tab<true>::EXP_N] = PICARD_EXP_TAB_BIG_INIT;
static __device__ const double g_log_tab_big[dmath::Tab<true>::LOG_N] = PICARD_LOG_TAB_BIG_INIT;

// stage a table set into shared memory ([exp][log]); the log part only when a log-likelihood is evaluated
template <bool BIG>
__device__ __forceinline__ void load_density_tables(double* tab, bool need_log, int tid, int nthreads) {
  using TB = dmath::Tab<BIG>;
  const double* ge = BIG ? g_exp_tab_big : g_exp_tab;
  const double* gl = BIG ? g_log_tab_big : g_log_tab;
  for (int i = tid; i < TB::EXP_N; i += nthreads) tab[i] = ge[i];
  if (need_log)
    for (int i = tid; i < TB::LOG_N; i += nthreads) tab[TB::EXP_N + i] = gl[i];
}

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA: 2-D tiled tensor load, completion by mbarrier transaction bytes (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// Decoupled stage recycling: every warp calls this once per tile after its last read of the stage; the warp that
// arrives last (shared-memory counter) runs `refill` (re-arms the mbarrier and issues the next TMA load).
template <int NWARPS, typename F>
__device__ __forceinline__ void stage_release(int* counter, int lane, F&& refill) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();                       // this warp's shared-memory reads of the stage are done (measured: free)
    const int old = atomicAdd(counter, 1);
    if (old == NWARPS - 1) {
      atomicExch(counter, 0);
      __threadfence_block();
      fence_proxy_async();                       // generic-proxy reads before the async-proxy (TMA) overwrite
      refill();
    }
  }
}
// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
}  // namespace ptx

template <int NP, int DENS, int MODE, bool WANT_H>
__global__ void __launch_bounds__(PassGeom<NP>::NTHREADS, PassGeom<NP>::MIN_BLOCKS)
pass_kernel(const __grid_constant__ CUtensorMap tmap, const PassParams p) {
  using G = PassGeom<NP>;
  constexpr bool WANT_G = (MODE == PASS_FUSED || MODE == PASS_GRAD);
  constexpr bool WANT_L = (MODE == PASS_FUSED || MODE == PASS_LOSS);
  constexpr bool APPLY = (MODE == PASS_APPLY);
  constexpr bool WANT_HM = WANT_G && WANT_H;  // the H matrix; in LOSS mode WANT_H only asks for the row sums of y^2
  constexpr bool WANT_SQ = WANT_H;            // Sq is used by the non-ortho Hessian / loss only (core.rs:80-81, 274)
  constexpr bool HAS_BIAS = APPLY || DENS == DENS_LINEAR;  // centering folded into the pass: never on the core-loop path
  constexpr bool NEED_TAB = !APPLY && (DENS == DENS_TANH || DENS == DENS_EXP);
  constexpr int MB = G::MB, NB = G::NB, WP = G::WPITCH, YP = G::YPITCH;

  // 1024-byte alignment (SWIZZLE_128B atom) comes from the declaration: keeping the pointer arithmetic free of integer
  // casts keeps every access in the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* ws = xs + G::STAGES * NP * G::BT;
  double* ys = ws + NP * WP;
  double* bs = ys + 2 * NP * YP;
  double* tab = bs + NP;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tab + dmath::TAB_DOUBLES);
  int* cnt = reinterpret_cast<int*>(bar + G::STAGES);  // per-stage release counters (LOSS mode: warps run decoupled)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = lane & 3, c = lane >> 2;

  // ---- W -> shared memory, zero-padded to NP x NP, columns permuted so that the A fragment of
  // k-step s is 4 contiguous doubles: column kk lives at 8*(kk/8) + 4*(kk%2) + (kk%8)/2.
  for (int idx = tid; idx < NP * NP; idx += G::NTHREADS) {
    int i = idx / NP, kk = idx % NP;
    double v = (i < p.n_out && kk < p.n_in) ? p.w[(size_t)i * p.ldw + kk] : 0.0;
    ws[i * WP + 8 * (kk >> 3) + 4 * (kk & 1) + ((kk & 7) >> 1)] = v;
  }
  for (int i = tid; i < NP; i += G::NTHREADS) bs[i] = (p.bias != nullptr && i < p.n_out) ? p.bias[i] : 0.0;
  if (NEED_TAB)
    load_density_tables<false>(tab, DENS == DENS_TANH, tid, G::NTHREADS);
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int s = 0; s < G::STAGES; ++s) { ptx::mbar_init(&bar[s], 1); cnt[s] = 0; }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  const int64_t tile0 = blockIdx.x, tstride = gridDim.x;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  constexpr uint32_t STAGE_BYTES = NP * G::BT * 8;
  if (tid == 0) {
    for (int s = 0; s < G::STAGES && s < my_tiles; ++s) {
      ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
      ptx::tma_load_2d(xs + s * NP * G::BT, &tmap, (int)((tile0 + s * tstride) * G::BT), 0, &bar[s]);
    }
  }

  // ---- persistent accumulators
  double gacc[WANT_G ? MB : 1][WANT_G ? NB : 1][2];
  double hacc[WANT_HM ? MB : 1][WANT_HM ? NB : 1][2];
#pragma unroll
  for (int a = 0; a < (WANT_G ? MB : 1); ++a)
#pragma unroll
    for (int b = 0; b < (WANT_G ? NB : 1); ++b) gacc[a][b][0] = gacc[a][b][1] = 0.0;
#pragma unroll
  for (int a = 0; a < (WANT_HM ? MB : 1); ++a)
#pragma unroll
    for (int b = 0; b < (WANT_HM ? NB : 1); ++b) hacc[a][b][0] = hacc[a][b][1] = 0.0;
  double sd[MB], sq[MB], sl[MB], brow[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    sd[mb] = sq[mb] = sl[mb] = 0.0;
    brow[mb] = bs[8 * (MB * warp + mb) + c];
  }

  // lane-constant shared-memory offsets (in doubles)
  // step-1 B fragment: channel k = 8*(s/2) + 2j + (s&1), sample 8nb + c; swizzle: 16-B chunk ^= (k & 7)
  int xoff[2][2];
#pragma unroll
  for (int nb = 0; nb < 2; ++nb)
#pragma unroll
    for (int b = 0; b < 2; ++b) xoff[nb][b] = (2 * j + b) * G::BT + (((4 * nb + (c >> 1)) ^ (2 * j + b)) << 1) + (c & 1);
  const int woff = (8 * MB * warp + c) * WP + j;  // + mb*8*WP + 4*s
  const int yoff_st = (8 * MB * warp + c) * YP + j;  // store: + mb*8*YP + 8nb + 4pp
  const int yoff_ld = c * YP + j;                    // load : + nbg*8*YP + 8nb' + 4pp

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = (int)(it % G::STAGES);
    const uint32_t parity = (uint32_t)((it / G::STAGES) & 1);
    const int64_t t0 = (tile0 + it * tstride) * G::BT;
    const bool partial_tile = (t0 + G::BT > p.t_local);
    const double* xst = xs + stage * NP * G::BT;
    double* yst = ys + (int)(it & 1) * NP * YP;

    ptx::mbar_wait(&bar[stage], parity);

    // ---------------- step 1: Y tile = W X (DMMA) ----------------
    double acc[MB][2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NP / 4; ++s) {
      double a[MB], b[2];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) a[mb] = ws[woff + mb * 8 * WP + 4 * s];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) b[nb] = xst[(s >> 1) * 8 * G::BT + xoff[nb][s & 1]];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) ptx::dmma(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
    }

    // ---------------- step 2: densities on the accumulator fragments ----------------
    double psi[WANT_G ? MB : 1][2][2], psd[WANT_HM ? MB : 1][2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          double y = HAS_BIAS ? acc[mb][nb][pp] - brow[mb] : acc[mb][nb][pp];
          const int64_t t = t0 + 8 * nb + 2 * j + pp;
          const bool valid = !partial_tile || (t < p.t_local);
          if (partial_tile && !valid) y = 0.0;  // zero-filled columns: harmless for Gr, Hr, Sq (every term has a factor y)
          if (APPLY) {
            acc[mb][nb][pp] = y;
          } else {
            double f = 0.0, fd = 0.0, dsd = 0.0, dsl = 0.0;
            density_eval<DENS, WANT_G, WANT_L>(y, p.dp, tab, f, fd, dsd, dsl);
            if (valid) {  // psi'(0) and loglik(0) are non-zero: padding columns must not reach Sd and L
              if (WANT_G) sd[mb] += dsd;
              if (WANT_L) sl[mb] += dsl;
            }
            if (WANT_G) psi[mb][nb][pp] = f;
            if (WANT_HM) psd[mb][nb][pp] = fd;
            if (WANT_SQ) sq[mb] = fma(y, y, sq[mb]);
            if (WANT_G) yst[yoff_st + mb * 8 * YP + 8 * nb + 4 * pp] = y;
          }
        }
    if (APPLY || (MODE == PASS_LOSS && p.out != nullptr)) {  // LOSS + out: keep Y' so an accepted try needs no W X product again
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int row = 8 * (MB * warp + mb) + c;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const int64_t t = t0 + 8 * nb + 2 * j;
          if (row < p.n_out) {
            double* dst = p.out + (size_t)row * p.ld_out + t;
            if (t + 1 < p.t_local) *reinterpret_cast<double2*>(dst) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            else if (t < p.t_local) dst[0] = acc[mb][nb][0];
          }
        }
      }
    }

    if (MODE == PASS_LOSS || APPLY) {
      // no data is exchanged between warps in these modes: no CTA barrier.  Warps drift apart, so one warp's density
      // evaluation overlaps another's DMMA phase on the shared FP64 pipe; the LAST warp to finish with a stage refills it.
      ptx::stage_release<G::NWARPS>(&cnt[stage], lane, [&] {
        if (it + G::STAGES < my_tiles) {
          ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
          ptx::tma_load_2d(xs + stage * NP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
        }
      });
    } else {
      __syncthreads();  // Y tile complete in shared memory; every warp is done with this X stage
      if (tid == 0 && it + G::STAGES < my_tiles) {
        ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
        ptx::tma_load_2d(xs + stage * NP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
      }
    }

    // ---------------- step 3: G += psi(Y) Y^T, H += psi'(Y) (Y^2)^T (DMMA) ----------------
    if (WANT_G) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // k-step = (nb', pp): 4 samples 8nb' + 2j' + pp, j' = 0..3
        const int nbp = ks >> 1, pp = ks & 1;
#pragma unroll
        for (int nbg = 0; nbg < NB; ++nbg) {
          const double b = yst[yoff_ld + nbg * 8 * YP + 8 * nbp + 4 * pp];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) ptx::dmma(gacc[mb][nbg][0], gacc[mb][nbg][1], psi[mb][nbp][pp], b);
          if (WANT_HM) {
            const double b2 = b * b;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) ptx::dmma(hacc[mb][nbg][0], hacc[mb][nbg][1], psd[mb][nbp][pp], b2);
          }
        }
      }
    }
  }

  // ---------------- per-CTA partial ----------------
  if (!APPLY) {
    double* part = p.partial + (size_t)blockIdx.x * pass_partial_size(NP, WANT_G, WANT_HM);
    if (WANT_G) {
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nbg = 0; nbg < NB; ++nbg) {
          const int row = 8 * (MB * warp + mb) + c, col = 8 * nbg + 2 * j;
          *reinterpret_cast<double2*>(part + row * NP + col) = make_double2(gacc[mb][nbg][0], gacc[mb][nbg][1]);
          if (WANT_HM)
            *reinterpret_cast<double2*>(part + NP * NP + row * NP + col) = make_double2(hacc[mb][nbg][0], hacc[mb][nbg][1]);
        }
    }
    double* rs = part + (WANT_G ? NP * NP : 0) + (WANT_HM ? NP * NP : 0);
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      double a = sd[mb], b = sq[mb], l = sl[mb];
      a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
      b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
      l += __shfl_xor_sync(0xffffffffu, l, 1); l += __shfl_xor_sync(0xffffffffu, l, 2);
      if (j == 0) {
        const int row = 8 * (MB * warp + mb) + c;
        rs[row] = a; rs[NP + row] = b; rs[2 * NP + row] = l;
      }
    }
  }
}

// ---- host-side launch description
struct PassLaunch {
  const double* d_x;     // (n_in x t_local), leading dimension ldx (even), device, 16-byte aligned
  int64_t ldx;
  int64_t t_local;
  int n_in, n_out;
  const double* d_w; int ldw;   // device
  const double* d_bias;         // device or nullptr
  int dens; double alpha;
  int mode; bool want_h;
  double* d_partial;            // workspace, >= max_grid * partial size
  double* d_mom;                // compact moments (moments modes)
  double* d_out; int64_t ld_out;  // APPLY
  int sm_count;
  cudaStream_t stream;
};

// Returns the number of kernels launched.
int launch_pass(const PassLaunch& L);
// TMA tensor map over a (n_rows x t_local) f64 matrix, box = np rows x 16 samples, SWIZZLE_128B, OOB zero fill
CUtensorMap make_tmap(const double* d_x, int64_t ldx, int64_t t_local, int n_rows, int np);
CUtensorMap make_tmap_box(const double* d_x, int64_t ldx, int64_t t_local, int n_rows, int box_cols, int box_rows, bool swizzle128);
int pass_padded_size(int n);  // NP for n (throws if unsupported)
size_t pass_workspace_doubles(int n, int sm_count);  // partial workspace needed for any mode

template <int NP>
int launch_pass_np(const PassLaunch& L, const CUtensorMap& tmap);
// row-block kernels (rowblock.cuh): LOSS / APPLY with W fragments in registers (KP = 128, 256) and the stored-Y gradient
template <int KP>
int launch_rb_loss(const PassLaunch& L, const CUtensorMap& tmap);
template <int NP>
int launch_rb_grady(const PassLaunch& L, const CUtensorMap& tmap);

}  // namespace picard
