// Launchers of the row-block kernels (rowblock.cuh); one translation unit per padded size (rb_np*.cu).
#pragma once
#include "rowblock.cuh"

namespace picard {

template <typename K>
static int rb_blocks_per_sm(K kern, int threads, size_t smem) {
  PICARD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int b = 0;
  PICARD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem));
  if (b < 1) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: row-block kernel does not fit on this device");
  return b > PASS_MAX_BLOCKS_PER_SM ? PASS_MAX_BLOCKS_PER_SM : b;
}

static inline int rb_reduce(const PassLaunch& L, int n_tg, int nrb, int rp, int np, bool wg, bool wh, bool wl) {
  const int n = L.n_out;
  const int64_t total = (wg ? (int64_t)n * n : 0) + (wh ? (int64_t)n * n : 0) + 3 * (int64_t)n;
  int rbk = (int)((total + 255) / 256);
  if (rbk > 4 * L.sm_count) rbk = 4 * L.sm_count;
  reduce_rb_kernel<<<rbk, 256, 0, L.stream>>>(L.d_partial, n_tg, nrb, rp, np, n, wg ? 1 : 0, wh ? 1 : 0, wl ? 1 : 0, L.d_mom);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}

template <int KP, int DENS, int MODE, bool WANT_SQ>
static int launch_rb_loss_one(const PassLaunch& L, const CUtensorMap& tmap) {
  using G = RbLossGeom<KP>;
  auto kern = rb_loss_kernel<KP, DENS, MODE, WANT_SQ>;
  static PerDeviceInt cache;  // per instantiation and per device
  const int bps = cache.get([&] { return rb_blocks_per_sm(kern, G::NTHREADS, G::SMEM_BYTES); });
  const int64_t n_tiles = (L.t_local + G::BT - 1) / G::BT;
  const int nrb = (L.n_out + G::RP - 1) / G::RP;
  int64_t n_tg = ((int64_t)L.sm_count * bps) / nrb;
  if (n_tg > n_tiles) n_tg = n_tiles;
  if (n_tg < 1) n_tg = 1;
  PassParams p;
  p.w = L.d_w; p.bias = L.d_bias; p.n_out = L.n_out; p.n_in = L.n_in; p.ldw = L.ldw;
  p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.partial = L.d_partial; p.out = L.d_out; p.ld_out = L.ld_out;
  kern<<<(unsigned)(n_tg * nrb), G::NTHREADS, G::SMEM_BYTES, L.stream>>>(tmap, p, nrb);
  PICARD_CUDA(cudaGetLastError());
  int launches = 1;
  if (MODE != PASS_APPLY) launches += rb_reduce(L, (int)n_tg, nrb, G::RP, KP, false, false, true);
  return launches;
}

template <int KP, int DENS>
static int launch_rb_loss_dens(const PassLaunch& L, const CUtensorMap& tmap) {
  return L.want_h ? launch_rb_loss_one<KP, DENS, PASS_LOSS, true>(L, tmap) : launch_rb_loss_one<KP, DENS, PASS_LOSS, false>(L, tmap);
}

template <int KP>
int launch_rb_loss(const PassLaunch& L, const CUtensorMap& tmap) {
  if (L.mode == PASS_APPLY) return launch_rb_loss_one<KP, DENS_LINEAR, PASS_APPLY, false>(L, tmap);
  switch (L.dens) {
    case DENS_TANH: return launch_rb_loss_dens<KP, DENS_TANH>(L, tmap);
    case DENS_EXP: return launch_rb_loss_dens<KP, DENS_EXP>(L, tmap);
    case DENS_CUBE: return launch_rb_loss_dens<KP, DENS_CUBE>(L, tmap);
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad density for the loss pass");
}

template <int NP, int DENS, bool WANT_H, bool HAS_BIAS>
static int launch_rb_grady_one(const PassLaunch& L, const CUtensorMap& tmap) {
  using G = RbGradYGeom<NP, WANT_H>;
  auto kern = rb_grady_kernel<NP, DENS, WANT_H, HAS_BIAS>;
  static PerDeviceInt cache;  // per instantiation and per device
  const int bps = cache.get([&] { return rb_blocks_per_sm(kern, G::NTHREADS, G::SMEM_BYTES); });
  const int64_t n_tiles = (L.t_local + G::BT - 1) / G::BT;
  const int nrb = (L.n_out + G::RP - 1) / G::RP;
  int64_t n_tg = ((int64_t)L.sm_count * bps) / nrb;
  if (n_tg > n_tiles) n_tg = n_tiles;
  if (n_tg < 1) n_tg = 1;
  PassParams p;
  p.w = nullptr; p.bias = L.d_bias; p.n_out = L.n_out; p.n_in = L.n_in; p.ldw = 0;
  p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.partial = L.d_partial; p.out = nullptr; p.ld_out = 0;
  kern<<<(unsigned)(n_tg * nrb), G::NTHREADS, G::SMEM_BYTES, L.stream>>>(tmap, p, nrb);
  PICARD_CUDA(cudaGetLastError());
  return 1 + rb_reduce(L, (int)n_tg, nrb, G::RP, NP, true, WANT_H, false);
}

template <int NP>
int launch_rb_grady(const PassLaunch& L, const CUtensorMap& tmap) {
  switch (L.dens) {
    case DENS_TANH: return L.want_h ? launch_rb_grady_one<NP, DENS_TANH, true, false>(L, tmap) : launch_rb_grady_one<NP, DENS_TANH, false, false>(L, tmap);
    case DENS_EXP: return L.want_h ? launch_rb_grady_one<NP, DENS_EXP, true, false>(L, tmap) : launch_rb_grady_one<NP, DENS_EXP, false, false>(L, tmap);
    case DENS_CUBE: return L.want_h ? launch_rb_grady_one<NP, DENS_CUBE, true, false>(L, tmap) : launch_rb_grady_one<NP, DENS_CUBE, false, false>(L, tmap);
    case DENS_LINEAR:  // Gram matrix of (X - mean): the covariance pass of the whitening step, and C = Y Y^T / T
      if (!L.want_h) return launch_rb_grady_one<NP, DENS_LINEAR, false, true>(L, tmap);
      break;
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad density / mode combination for the stored-Y gradient pass");
}

}  // namespace picard
