// JADE warm start on the device (jade.rs:22-197): cumulant Gram kernel (K8) + persistent Jacobi sweeps (K9).
#pragma once
#include "engine.cuh"

namespace picard {
// d_x: whitened data (n x t_local, leading dimension ld) on the device; w_out (n x n, host) receives
// sym_decorrelation(V) (jade.rs:69).  sweeps_done may be NULL.
void jade_device(const double* d_x, int n, int64_t t_local, int64_t ld, double t_total, int64_t max_iter, double tol, bool verbose,
                 picard_comm* comm, int sm_count, cudaStream_t st, double* w_out, int64_t* sweeps_done, picard_stats_t* stats);
// cumulant matrices only (test hook): out is [n(n+1)/2][n][n] on the host
void jade_cumulants_device(const double* d_x, int n, int64_t t_local, int64_t ld, double t_total, picard_comm* comm, int sm_count,
                           cudaStream_t st, double* out_host, picard_stats_t* stats);
}  // namespace picard
