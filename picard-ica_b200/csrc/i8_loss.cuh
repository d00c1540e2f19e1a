// INT8 tensor-core form of the LOSS pass (i8_loss.cu): error-free 7-slice splitting, tcgen05.mma kind::i8, accumulators in TMEM.
// 64 < N <= 128, whitened data (PICARD_I8=0 / 1 overrides); see the header of i8_loss.cu.
#pragma once
#include "pass.cuh"

namespace picard {

constexpr int I8_SLICES = 7;                                     // signed 7-bit slices per operand
constexpr int I8_TILE = 32;                                      // samples per tile
constexpr int I8_TILE_BYTES = I8_SLICES * I8_TILE * 128 + I8_TILE * 8;   // 7 slices of [32 samples][128 bytes] + 32 column scales
constexpr int I8_WBLOB_BYTES = I8_SLICES * 128 * 128 + 128 * 8;          // 7 slices of W' + 128 row scales

int i8_mode();                           // PICARD_I8: 1 = force, 0 = off, unset (-1) = automatic (whitened data only)
size_t i8_blob_bytes(int64_t t_local);   // size of the sliced image of an (n <= 128) x t_local sample matrix
// x (n x t_local, leading dimension ldx) -> sliced tiles; once per fit (x1 does not change during the core loop)
int i8_slice_x(const double* d_x, int64_t ldx, int64_t t_local, int n, uint8_t* blob, cudaStream_t st);
// LOSS pass at W = L.d_w from the sliced image; wblob: I8_WBLOB_BYTES of device scratch.  Returns the number of kernels launched.
int launch_loss_i8(const PassLaunch& L, const uint8_t* xblob, uint8_t* wblob);

}  // namespace picard
