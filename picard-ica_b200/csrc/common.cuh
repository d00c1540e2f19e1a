// Shared host/device helpers for libpicard_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <atomic>
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

// Per-device, thread-safe cache of a launch constant (blocks per SM after cudaFuncSetAttribute, or just "configured").
// Function attributes and occupancy are properties of (kernel, device): a process-wide static would leave the kernel
// unconfigured on the second device a process fits on.  compute() is idempotent, so a benign race only repeats it.
struct PerDeviceInt {
  static constexpr int kMaxDevices = 64;
  std::atomic<int> v[kMaxDevices] = {};  // 0 = not computed yet on that device
  template <typename F>
  int get(F&& compute) {
    int dev = 0;
    PICARD_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return compute();
    int x = v[dev].load(std::memory_order_acquire);
    if (x == 0) { x = compute(); v[dev].store(x, std::memory_order_release); }
    return x;
  }
};

// Density kinds on the device. LINEAR (psi(y) = y) is internal: it turns the moments pass into the
// covariance SYRK of the whitening step (whitening.rs:61 replacement).
enum : int { DENS_TANH = 0, DENS_EXP = 1, DENS_CUBE = 2, DENS_LINEAR = 3 };

// Pass modes
enum : int { PASS_FUSED = 0, PASS_GRAD = 1, PASS_LOSS = 2, PASS_APPLY = 3, PASS_GRADY = 4 /* gradient moments from a stored Y */ };

}  // namespace picard
