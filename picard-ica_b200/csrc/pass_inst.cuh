// Instantiation body shared by pass_np*.cu: one translation unit per padded size so the (slow, fully
// unrolled) kernels compile in parallel.
#pragma once
#include "rowblock.cuh"

namespace picard {

template <int NP, int DENS, int MODE, bool WANT_H>
static int launch_one(const PassLaunch& L, const CUtensorMap& tmap) {
  using G = PassGeom<NP>;
  auto kern = pass_kernel<NP, DENS, MODE, WANT_H>;
  static PerDeviceInt cache;  // per instantiation and per device
  const int blocks_per_sm = cache.get([&] {
    PICARD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    int b = 0;
    PICARD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, G::NTHREADS, G::SMEM_BYTES));
    if (b < 1) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: pass kernel does not fit on this device");
    return b > PASS_MAX_BLOCKS_PER_SM ? PASS_MAX_BLOCKS_PER_SM : b;
  });
  const int64_t n_tiles = (L.t_local + G::BT - 1) / G::BT;
  int64_t grid = (int64_t)L.sm_count * blocks_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  PassParams p;
  p.w = L.d_w; p.bias = L.d_bias; p.n_out = L.n_out; p.n_in = L.n_in; p.ldw = L.ldw;
  p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.partial = L.d_partial; p.out = L.d_out; p.ld_out = L.ld_out;
  kern<<<(unsigned)grid, G::NTHREADS, G::SMEM_BYTES, L.stream>>>(tmap, p);
  PICARD_CUDA(cudaGetLastError());
  int launches = 1;
  if (MODE != PASS_APPLY) {
    constexpr bool WG = (MODE == PASS_FUSED || MODE == PASS_GRAD), WL = (MODE == PASS_FUSED || MODE == PASS_LOSS);
    const int n = L.n_out;
    constexpr bool WHM = WG && WANT_H;  // the H matrix (in LOSS mode WANT_H only asks for Sq)
    const int64_t total = (WG ? (int64_t)n * n : 0) + (WHM ? (int64_t)n * n : 0) + 3 * (int64_t)n;
    int rb = (int)((total + 255) / 256);
    if (rb > 4 * L.sm_count) rb = 4 * L.sm_count;
    reduce_rb_kernel<<<rb, 256, 0, L.stream>>>(L.d_partial, (int)grid, 1, NP, NP, n, WG ? 1 : 0, WHM ? 1 : 0, WL ? 1 : 0, L.d_mom);
    PICARD_CUDA(cudaGetLastError());
    ++launches;
  }
  return launches;
}

template <int NP, int DENS>
static int launch_dens(const PassLaunch& L, const CUtensorMap& tmap) {
  switch (L.mode) {
    case PASS_FUSED: return L.want_h ? launch_one<NP, DENS, PASS_FUSED, true>(L, tmap) : launch_one<NP, DENS, PASS_FUSED, false>(L, tmap);
    case PASS_GRAD: return L.want_h ? launch_one<NP, DENS, PASS_GRAD, true>(L, tmap) : launch_one<NP, DENS, PASS_GRAD, false>(L, tmap);
    case PASS_LOSS: return L.want_h ? launch_one<NP, DENS, PASS_LOSS, true>(L, tmap) : launch_one<NP, DENS, PASS_LOSS, false>(L, tmap);
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad pass mode");
}

template <int NP>
int launch_pass_np(const PassLaunch& L, const CUtensorMap& tmap) {
  if (L.mode == PASS_APPLY) return launch_one<NP, DENS_LINEAR, PASS_APPLY, false>(L, tmap);
  switch (L.dens) {
    case DENS_TANH: return launch_dens<NP, DENS_TANH>(L, tmap);
    case DENS_EXP: return launch_dens<NP, DENS_EXP>(L, tmap);
    case DENS_CUBE: return launch_dens<NP, DENS_CUBE>(L, tmap);
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad density / mode combination");
}

}  // namespace picard
