// Host side of the INT8 LOSS pass (kernels: i8_loss_kernel.cuh; numerics: i8_common.cuh).
#include "i8_loss_kernel.cuh"

#include "rowblock_inst.cuh"

namespace picard {

int i8_env_mode() {
  static const int v = [] {
    const char* e = getenv("PICARD_I8");
    return !e ? -1 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : -1));
  }();
  return v;
}

size_t i8_blob_bytes(int64_t t_local) { return (size_t)((t_local + I8_TILE - 1) / I8_TILE) * i8::LossGeom<I8_TILE>::TILE_BYTES; }

int i8_slice_x(const double* d_x, int64_t ldx, int64_t t_local, int n, uint8_t* blob, double* d_stats, int sm_count, cudaStream_t st) {
  const int64_t n_tiles = (t_local + I8_TILE - 1) / I8_TILE;
  int64_t grid = (int64_t)sm_count * 8;
  if (grid > n_tiles) grid = n_tiles;
  PICARD_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(double) * I8_XSTATS, st));
  i8::slice_x_kernel<I8_TILE><<<(unsigned)grid, 8 * I8_TILE, 0, st>>>(d_x, ldx, t_local, n, n_tiles, blob, d_stats);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}

template <int DENS, bool WANT_SQ>
static int launch_loss_i8_one(const PassLaunch& L, const uint8_t* xblob, const I8LossFinish& fin) {
  using G = i8::LossGeom<I8_TILE>;
  auto kern = i8::loss_i8_kernel<DENS, WANT_SQ, I8_TILE, 0>;
  if (L.d_out != nullptr && ((L.ld_out % I8_TILE) != 0 || (reinterpret_cast<uintptr_t>(L.d_out) & 15) != 0))
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the Y store of the INT8 LOSS pass needs a 16-byte aligned base and a leading dimension that is a multiple of the tile");
  static PerDeviceInt configured;  // per instantiation and per device
  configured.get([&] {
    PICARD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    return 1;
  });
  const int64_t n_tiles = (L.t_local + I8_TILE - 1) / I8_TILE;
  int64_t grid = L.sm_count;
  if (grid > n_tiles) grid = n_tiles;
  PassParams p;
  p.w = L.d_w; p.bias = nullptr; p.n_out = L.n_out; p.n_in = L.n_in; p.ldw = L.ldw;
  p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.partial = L.d_partial; p.out = L.d_out; p.ld_out = L.ld_out;
  i8::LossTail tail;
  if (fin.counter != nullptr) {
    *fin.counter_total += (unsigned int)grid;
    tail.counter = fin.counter; tail.target = *fin.counter_total; tail.mom = L.d_mom; tail.finish = fin.finish; tail.which = fin.which;
    if (fin.dims) tail.dims = *static_cast<const CoreDims*>(fin.dims);
    if (fin.px) { tail.exchange = 1; tail.px = *static_cast<const P2PCall*>(fin.px); }
    tail.signs = fin.signs; tail.sc = static_cast<CoreScalars*>(fin.sc); tail.sc_map = static_cast<CoreScalars*>(fin.sc_map); tail.seq = fin.seq;
  }
  const CUtensorMap tmap_out = L.d_out != nullptr ? make_tmap_box(L.d_out, L.ld_out, L.t_local, L.n_out, G::CPT, 32, true) : CUtensorMap{};
  kern<<<(unsigned)grid, G::NTHREADS, G::SMEM_BYTES, L.stream>>>(xblob, tmap_out, p, tail, nullptr);
  PICARD_CUDA(cudaGetLastError());
  if (fin.counter != nullptr) return 1;
  return 1 + rb_reduce(L, (int)grid, 1, 128, 128, false, false, true);
}

int launch_loss_i8(const PassLaunch& L, const uint8_t* xblob, const I8LossFinish& fin) {
  if (L.mode != PASS_LOSS || L.n_in > 128 || L.n_out > 128 || L.d_bias != nullptr)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the INT8 pass covers the LOSS mode for N <= 128 only");
  switch (L.dens) {
    case DENS_TANH: return L.want_h ? launch_loss_i8_one<DENS_TANH, true>(L, xblob, fin) : launch_loss_i8_one<DENS_TANH, false>(L, xblob, fin);
    case DENS_EXP: return L.want_h ? launch_loss_i8_one<DENS_EXP, true>(L, xblob, fin) : launch_loss_i8_one<DENS_EXP, false>(L, xblob, fin);
    case DENS_CUBE: return L.want_h ? launch_loss_i8_one<DENS_CUBE, true>(L, xblob, fin) : launch_loss_i8_one<DENS_CUBE, false>(L, xblob, fin);
    default: break;
  }
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: bad density for the INT8 loss pass");
}

}  // namespace picard
