// INT8 tensor-core forms of the two passes of the core loop (tcgen05.mma kind::i8, s32 accumulators in TMEM), whitened problems
// with 64 < N <= 128:
//   i8_loss.cu : the LOSS pass of a line-search try (Y' = W' x1, log-likelihood / y^2 row sums, Y' kept in HBM)
//   i8_grad.cu : the stored-Y gradient pass (Gr = psi(Y) Y^T, Sd)
// Both rest on the error-free balanced radix-256 splitting of i8_common.cuh.
#pragma once
#include "pass.cuh"

namespace picard {

constexpr int I8_SLICES = 6;   // balanced radix-256 digits per operand
constexpr int I8_TILE = 64;    // samples per tile of the sliced image of x1
// statistics of x1 gathered while it is sliced: [0] sum_t 2^(e_t - 1) (the power-of-two bounds of the samples' largest components),
// [1] max_t |x_t|^2 as the bit pattern of a double, [2 .. 2 + 128) row sums of squares
constexpr int I8_XSTATS = 2 + 128;

int i8_env_mode();                       // PICARD_I8: 1 = force, 0 = off, unset (-1) = automatic
size_t i8_blob_bytes(int64_t t_local);   // size of the sliced image of an (n <= 128) x t_local sample matrix
// x (n x t_local, leading dimension ldx) -> sliced tiles + statistics (d_stats: I8_XSTATS doubles, zeroed by this call); once
// per fit (x1 does not change during the core loop).  Returns the number of kernels launched.
int i8_slice_x(const double* d_x, int64_t ldx, int64_t t_local, int n, uint8_t* blob, double* d_stats, int sm_count, cudaStream_t st);
// What the LOSS pass does after its streaming part (see LossTail in i8_loss_kernel.cuh).  counter: one device unsigned int, zeroed
// once by the owner; *counter_total is the owner's running total of CTAs counted so far (host side).
struct I8LossFinish {
  unsigned int* counter = nullptr;
  unsigned int* counter_total = nullptr;
  int finish = 0;                 // 1: compute the loss / accept flag in the kernel and publish (single GPU)
  int which = 0;
  const void* dims = nullptr;     // const CoreDims*
  const double* signs = nullptr;
  void* sc = nullptr;             // CoreScalars* (device)
  void* sc_map = nullptr;         // CoreScalars* (pinned host mirror)
  unsigned long long seq = 0;
  const void* px = nullptr;       // const P2PCall*: exchange the row sums between the ranks inside the kernel (p2p.cuh)
};
// LOSS pass at W = L.d_w from the sliced image: one kernel (W' digits, streaming, reduction of the partials into L.d_mom [, loss]).
// Returns the number of kernels launched.
int launch_loss_i8(const PassLaunch& L, const uint8_t* xblob, const I8LossFinish& fin);

// Gradient moments Gr = psi(Y) Y^T and Sd = sum psi'(Y) from the stored Y (L.d_x = Y, n_in = n_out = N <= 128), ortho problems
// (no Hr), tanh / exp densities.  d_rowexp: N ints, e_j with 1.008 |y_jt| < 2^e_j for every t of this shard (i8_row_exponents).
bool i8_grad_supported(int n, int dens, bool want_h);
// counter / counter_total as in I8LossFinish (nullptr: partials + a separate reduction launch)
// px: const P2PCall* (exchange [Gr | Sd] between the ranks inside the kernel); counter has three entries then (see CoreSolver)
int launch_grad_i8(const PassLaunch& L, const int* d_rowexp, unsigned int* counter = nullptr, unsigned int* counter_total = nullptr,
                   const void* px = nullptr);
// e_j = bound_exponent(|w_j|_2 * xnorm_max) for the rows of W (n x n): the rigorous bound |y_jt| <= |w_j| |x_t|
int i8_row_exponents(const double* d_w, int n, const double* d_xstats, int* d_rowexp, cudaStream_t st);

}  // namespace picard
