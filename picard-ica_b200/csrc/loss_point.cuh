// The loss of a point from its reduced moments (core.rs:39-85 after the sums), one warp: shared by loss_kernel (small.cu) and the
// fused tail of the INT8 LOSS pass (i8_loss_kernel.cuh).
#pragma once
#include "pass.cuh"
#include "small.cuh"

namespace picard {
namespace small {

// one warp: lane-strided partial sums + shuffle tree (fixed order, deterministic)
__device__ __forceinline__ double loss_of_point_warp(const CoreDims& d, const double* mom, const double* signs, bool* singular) {
  const int n = d.n, lane = threadIdx.x & 31;
  const double tf = d.t_total;
  *singular = false;
  double base = 0.0;
  if (!d.ortho) {
    const double* ex = mom + mom_size(n);
    if (ex[1] == 0.0) { *singular = true; return 1e15; }
    base = -ex[0];
  }
  const double* L = mom + mom_off_ll(n);
  const double* Sq = mom + mom_off_sq(n);
  double part = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double s = signs ? signs[i] : 1.0;
    part += s * L[i] / tf;
    if (d.extended && !d.ortho) part += 0.5 * Sq[i] / tf;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return base + part;
}


}  // namespace small
}  // namespace picard
