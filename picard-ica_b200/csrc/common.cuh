// Shared host/device helpers for libpicard_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/picard_b200.h"

namespace picard {

// A status + message carried up to the C ABI (maps onto PicardError, error.rs:9-42).
struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

#define PICARD_CUDA(expr)                                                                                         \
  do {                                                                                                            \
    cudaError_t e__ = (expr);                                                                                     \
    if (e__ != cudaSuccess)                                                                                       \
      throw ::picard::Error(PICARD_COMPUTATION_ERROR, std::string("Computation error: CUDA failure '") +          \
                                                          cudaGetErrorString(e__) + "' in " #expr " (" __FILE__   \
                                                          ":" + std::to_string(__LINE__) + ")");                  \
  } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Density kinds on the device. LINEAR (psi(y) = y) is internal: it turns the moments pass into the
// covariance SYRK of the whitening step (whitening.rs:61 replacement).
enum : int { DENS_TANH = 0, DENS_EXP = 1, DENS_CUBE = 2, DENS_LINEAR = 3 };

// Pass modes
enum : int { PASS_FUSED = 0, PASS_GRAD = 1, PASS_LOSS = 2, PASS_APPLY = 3, PASS_GRADY = 4 /* gradient moments from a stored Y */ };

}  // namespace picard
