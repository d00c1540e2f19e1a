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
constexpr int I8_WBLOB_BYTES = I8_SLICES * 128 * 128 + 128 * 4;  // six slices of W' (row-major, 128 x 128 bytes) + 128 row exponents
// statistics of x1 gathered while it is sliced: [0] sum_t 2^(e_t - 1) (the power-of-two bounds of the samples' largest components),
// [1] max_t |x_t|^2 as the bit pattern of a double, [2 .. 2 + 128) row sums of squares
constexpr int I8_XSTATS = 2 + 128;

int i8_env_mode();                       // PICARD_I8: 1 = force, 0 = off, unset (-1) = automatic
size_t i8_blob_bytes(int64_t t_local);   // size of the sliced image of an (n <= 128) x t_local sample matrix
// x (n x t_local, leading dimension ldx) -> sliced tiles + statistics (d_stats: I8_XSTATS doubles, zeroed by this call); once
// per fit (x1 does not change during the core loop).  Returns the number of kernels launched.
int i8_slice_x(const double* d_x, int64_t ldx, int64_t t_local, int n, uint8_t* blob, double* d_stats, int sm_count, cudaStream_t st);
// LOSS pass at W = L.d_w from the sliced image; wblob: I8_WBLOB_BYTES of device scratch.  Returns the number of kernels launched.
int launch_loss_i8(const PassLaunch& L, const uint8_t* xblob, uint8_t* wblob);

// Gradient moments Gr = psi(Y) Y^T and Sd = sum psi'(Y) from the stored Y (L.d_x = Y, n_in = n_out = N <= 128), ortho problems
// (no Hr), tanh / exp densities.  d_rowexp: N ints, e_j with 1.008 |y_jt| < 2^e_j for every t of this shard (i8_row_exponents).
bool i8_grad_supported(int n, int dens, bool want_h);
int launch_grad_i8(const PassLaunch& L, const int* d_rowexp);
// e_j = bound_exponent(|w_j|_2 * xnorm_max) for the rows of W (n x n): the rigorous bound |y_jt| <= |w_j| |x_t|
int i8_row_exponents(const double* d_w, int n, const double* d_xstats, int* d_rowexp, cudaStream_t st);

}  // namespace picard
