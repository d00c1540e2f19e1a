// Host side of the INT8 gradient pass (kernel: i8_grad_kernel.cuh; numerics: i8_common.cuh).
#include "i8_grad_kernel.cuh"

#include "rowblock_inst.cuh"

namespace picard {

bool i8_grad_supported(int n, int dens, bool want_h) {
  return n > 64 && n <= 128 && !want_h && (dens == DENS_TANH || dens == DENS_EXP);
}

int i8_row_exponents(const double* d_w, int n, const double* d_xstats, int* d_rowexp, cudaStream_t st) {
  i8::row_exponent_kernel<<<(n + 7) / 8, 256, 0, st>>>(d_w, n, d_xstats, d_rowexp);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}

template <int DENS>
static int launch_grad_i8_one(const PassLaunch& L, const int* d_rowexp, unsigned int* counter, unsigned int* counter_total, const void* px) {
  using G = i8::GradGeom;
  auto kern = i8::grad_i8_kernel<DENS, 0, 1>;  // operand layout 1 (core matrices, no swizzle): 4 % faster than SWIZZLE_32B in profiles/lab
  static PerDeviceInt configured;  // per instantiation and per device
  configured.get([&] {
    PICARD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    return 1;
  });
  const CUtensorMap tmap = make_tmap(L.d_x, L.ldx, L.t_local, L.n_in, 128);
  const int64_t n_tiles = (L.t_local + G::KT - 1) / G::KT;
  int64_t n_tg = L.sm_count / 2;
  if (n_tg > n_tiles) n_tg = n_tiles;
  if (n_tg < 1) n_tg = 1;
  i8::GradParams p;
  p.n = L.n_out; p.t_local = L.t_local; p.n_tiles = n_tiles; p.dp = make_dens_params(DENS, L.alpha);
  p.rowexp = d_rowexp; p.psi_exp = i8::psi_exponent(DENS, L.alpha); p.partial = L.d_partial;
  p.counter = nullptr; p.target = 0; p.mom = L.d_mom; p.exchange = 0; p.counter2 = nullptr; p.target2 = 0; p.px = P2PCall{};
  if (counter != nullptr && 2 * n_tg <= L.sm_count) {
    // the tail's device-wide wait needs every CTA resident: cooperative launch (one CTA per SM, grid <= SM count)
    *counter_total += (unsigned int)(2 * n_tg);
    p.counter = counter; p.target = *counter_total;
    if (px) {  // counter[1] / counter_total[1]: the second device-wide count of the exchange
      counter_total[1] += (unsigned int)(2 * n_tg);
      p.exchange = 1; p.counter2 = counter + 1; p.target2 = counter_total[1]; p.px = *static_cast<const P2PCall*>(px);
    }
    long long* no_trace = nullptr;
    void* args[] = {(void*)&tmap, (void*)&p, (void*)&no_trace};
    PICARD_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(2 * n_tg)), dim3(G::NTHREADS), args, G::SMEM_BYTES, L.stream));
    return 1;
  }
  kern<<<(unsigned)(2 * n_tg), G::NTHREADS, G::SMEM_BYTES, L.stream>>>(tmap, p, nullptr);
  PICARD_CUDA(cudaGetLastError());
  return 1 + rb_reduce(L, (int)n_tg, 2, G::NB, G::MA, true, false, false);
}

int launch_grad_i8(const PassLaunch& L, const int* d_rowexp, unsigned int* counter, unsigned int* counter_total, const void* px) {
  if (L.mode != PASS_GRADY || !i8_grad_supported(L.n_out, L.dens, L.want_h) || L.n_in != L.n_out || L.d_bias != nullptr)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the INT8 gradient pass covers ortho problems with 64 < N <= 128, tanh / exp");
  return L.dens == DENS_TANH ? launch_grad_i8_one<DENS_TANH>(L, d_rowexp, counter, counter_total, px)
                             : launch_grad_i8_one<DENS_EXP>(L, d_rowexp, counter, counter_total, px);
}

}  // namespace picard
