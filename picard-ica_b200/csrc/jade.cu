// JADE warm start on the device (jade.rs:22-197).
//
//   K8  jade_cumulant_kernel   fourth-order cumulant sums  S_ij[k,l] = sum_s x_i x_j x_k x_l  for all pairs i <= j
//       (jade.rs:78-131 materialises x_i x_j as an N x N x T array and runs N(N+1)/2 * N^2 scalar dot products of
//       length T).  Here each S_ij = X diag(x_i . x_j) X^T is a weighted Gram matrix accumulated on DMMA from the
//       TMA-staged X tile: the tile's fragments are loaded once per warp and reused for every pair the warp owns;
//       the weight costs one DMUL per A fragment.  A CTA owns a group of pairs (their N x N accumulators live in
//       registers) and a share of the sample tiles; partials are reduced in a fixed order.
//       Algorithmic work: 2 * [N(N+1)/2] * N^2 * T flop (as the reference computes it), 8 N T bytes.
//   K9  jade_sweep_kernel      cyclic Jacobi sweeps in the reference's pair order (jade.rs:40-66).  The reference
//       re-evaluates the 2 x 2 block of V^T M V from scratch for every matrix and every pair (O(N^2) each); here the
//       rotated matrices M' = V^T M V are kept up to date (two rows + two columns per matrix per rotation, O(N)),
//       which is the same quantity in exact arithmetic.  Single CTA, matrices in L2 with the matrix index fastest so
//       that thread m <-> matrix m accesses coalesce.
//   then sym_decorrelation(V) (jade.rs:69; note: V, not V^T -- quirk Q17).
#include "jade.cuh"

#include <cmath>

namespace picard {

namespace {

template <int NP>
struct JadeGeom {
  static_assert(NP == 8 || NP == 16 || NP == 32, "JADE kernels are sized for N <= 32");
  static constexpr int NWARPS = 8;
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int NB = NP / 8;
  static constexpr int PPW = 2048 / (NP * NP) > 8 ? 8 : 2048 / (NP * NP);  // pairs per warp: <= 64 accumulator doubles per thread
  static constexpr int PPC = PPW * NWARPS;                                  // pairs per CTA
  static constexpr int BT = 16;
  static constexpr int STAGES = 4;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * NP * BT * 8 + 128;
};

struct JadeParams {
  int n, n_pairs;
  int64_t t_local, n_tiles;
  double* partial;  // [n_tg][n_pairs_padded][NP*NP]
  int n_pg;         // pair groups
};

__device__ __forceinline__ void pair_from_index(int m, int n, int& i, int& j) {  // m -> (i <= j), row-major upper triangle
  int ii = 0, rem = m;
  while (rem >= n - ii) { rem -= n - ii; ++ii; }
  i = ii; j = ii + rem;
}

template <int NP>
__global__ void __launch_bounds__(JadeGeom<NP>::NTHREADS, 1)
jade_cumulant_kernel(const __grid_constant__ CUtensorMap tmap, const JadeParams p) {
  using G = JadeGeom<NP>;
  constexpr int NB = G::NB, PPW = G::PPW;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(xs + G::STAGES * NP * G::BT);
  int* cnt = reinterpret_cast<int*>(bar + G::STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jj = lane & 3, c = lane >> 2;
  const int pg = blockIdx.x % p.n_pg, tg = blockIdx.x / p.n_pg, n_tg = gridDim.x / p.n_pg;
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int s = 0; s < G::STAGES; ++s) { ptx::mbar_init(&bar[s], 1); cnt[s] = 0; }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int64_t tile0 = tg, tstride = n_tg;
  const int64_t my_tiles = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  constexpr uint32_t STAGE_BYTES = NP * G::BT * 8;
  if (tid == 0) {
    for (int s = 0; s < G::STAGES && s < my_tiles; ++s) {
      ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
      ptx::tma_load_2d(xs + s * NP * G::BT, &tmap, (int)((tile0 + s * tstride) * G::BT), 0, &bar[s]);
    }
  }
  // this warp's pairs
  int pi[PPW], pj[PPW];
  bool live[PPW];
#pragma unroll
  for (int q = 0; q < PPW; ++q) {
    const int m = pg * G::PPC + warp * PPW + q;
    live[q] = m < p.n_pairs;
    pi[q] = 0; pj[q] = 0;
    if (live[q]) pair_from_index(m, p.n, pi[q], pj[q]);
  }
  double acc[PPW][NB][NB][2];
#pragma unroll
  for (int q = 0; q < PPW; ++q)
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[q][a][b][0] = acc[q][a][b][1] = 0.0;

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = (int)(it % G::STAGES);
    const uint32_t parity = (uint32_t)((it / G::STAGES) & 1);
    const double* xt = xs + stage * NP * G::BT;
    ptx::mbar_wait(&bar[stage], parity);
    // fragments of the tile, shared by the A and B operands: block blk, half nb' -> rows 8 blk + c, samples 2 (2 jj + nb') + {0, 1}
    double2 xf[NB][2];
#pragma unroll
    for (int blk = 0; blk < NB; ++blk)
#pragma unroll
      for (int nbp = 0; nbp < 2; ++nbp)
        xf[blk][nbp] = *reinterpret_cast<const double2*>(xt + (8 * blk + c) * G::BT + (((2 * jj + nbp) ^ c) << 1));
#pragma unroll
    for (int q = 0; q < PPW; ++q) {
      if (!live[q]) continue;  // warp-uniform
#pragma unroll
      for (int nbp = 0; nbp < 2; ++nbp) {
        // weights w[s] = x_i[s] x_j[s] for this lane's two samples of the half
        const double2 xi = *reinterpret_cast<const double2*>(xt + pi[q] * G::BT + (((2 * jj + nbp) ^ (pi[q] & 7)) << 1));
        const double2 xj = *reinterpret_cast<const double2*>(xt + pj[q] * G::BT + (((2 * jj + nbp) ^ (pj[q] & 7)) << 1));
        const double w0 = xi.x * xj.x, w1 = xi.y * xj.y;
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
          const double a0 = xf[mb][nbp].x * w0, a1 = xf[mb][nbp].y * w1;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            ptx::dmma(acc[q][mb][nb][0], acc[q][mb][nb][1], a0, xf[nb][nbp].x);
            ptx::dmma(acc[q][mb][nb][0], acc[q][mb][nb][1], a1, xf[nb][nbp].y);
          }
        }
      }
    }
    ptx::stage_release<G::NWARPS>(&cnt[stage], lane, [&] {
      if (it + G::STAGES < my_tiles) {
        ptx::mbar_expect_tx(&bar[stage], STAGE_BYTES);
        ptx::tma_load_2d(xs + stage * NP * G::BT, &tmap, (int)((tile0 + (it + G::STAGES) * tstride) * G::BT), 0, &bar[stage]);
      }
    });
  }
#pragma unroll
  for (int q = 0; q < PPW; ++q) {
    const int m = pg * G::PPC + warp * PPW + q;
    if (!live[q]) continue;
    double* dst = p.partial + ((size_t)tg * (size_t)(p.n_pg * G::PPC) + (size_t)m) * (NP * NP);
#pragma unroll
    for (int mb = 0; mb < NB; ++mb)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        *reinterpret_cast<double2*>(dst + (8 * mb + c) * NP + 8 * nb + 2 * jj) = make_double2(acc[q][mb][nb][0], acc[q][mb][nb][1]);
  }
}

// raw[m][k*n + l] (compact, ld n) = sum over tile groups of the partials (fixed order)
__global__ void jade_reduce_kernel(const double* __restrict__ partial, int n_tg, int pairs_padded, int np, int n, int n_pairs,
                                   double* __restrict__ raw) {
  const int64_t total = (int64_t)n_pairs * n * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e / (n * n)), r = (int)(e % (n * n)), k = r / n, l = r % n;
    double s = 0.0;
    for (int t = 0; t < n_tg; ++t) s += partial[((size_t)t * pairs_padded + m) * (np * np) + k * np + l];
    raw[e] = s;
  }
}

// jade.rs:101-127: Q_ij[k,l] = S_ij[k,l] / T - d_ij d_kl - d_ik d_jl - d_il d_jk, then (Q + Q^T) / 2.
// std_out: [m][k][l] (test hook, may be NULL); rot_out: [k][l][m_pad] (matrix index fastest, for the sweeps).
__global__ void jade_finalize_kernel(const double* __restrict__ raw, int n, int n_pairs, int m_pad, double t_total, double* std_out,
                                     double* rot_out) {
  const int64_t total = (int64_t)n_pairs * n * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e / (n * n)), r = (int)(e % (n * n)), k = r / n, l = r % n;
    int i, j;
    pair_from_index(m, n, i, j);
    auto q = [&](int kk, int ll) {
      double v = raw[(size_t)m * n * n + kk * n + ll] / t_total;
      if (i == j && kk == ll) v -= 1.0;
      if (i == kk && j == ll) v -= 1.0;
      if (i == ll && j == kk) v -= 1.0;
      return v;
    };
    const double v = (q(k, l) + q(l, k)) / 2.0;
    if (std_out) std_out[e] = v;
    if (rot_out) rot_out[((size_t)k * n + l) * m_pad + m] = v;
  }
}

// Jacobi sweeps (jade.rs:40-66, 137-197), reference pair order, single CTA.  Mr: [k][l][m_pad], rotated in place.
__global__ void __launch_bounds__(1024) jade_sweep_kernel(double* __restrict__ Mr, int n, int n_pairs, int m_pad, int max_iter, double tol,
                                                          double* __restrict__ V_out, int* __restrict__ sweeps_out) {
  extern __shared__ double vsm[];  // V (n x n)
  __shared__ double sh[3][33];
  __shared__ double s_cs[3];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int e = tid; e < n * n; e += nt) vsm[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();
  int sweeps = 0;
  for (int iter = 0; iter < max_iter; ++iter) {
    double max_theta = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        // ---- g accumulated over all matrices from the current 2 x 2 blocks of M' = V^T M V (jade.rs:146-167)
        double g00 = 0.0, g01 = 0.0, g11 = 0.0;
        for (int m = tid; m < n_pairs; m += nt) {
          const double bpp = Mr[((size_t)p * n + p) * m_pad + m], bpq = Mr[((size_t)p * n + q) * m_pad + m];
          const double bqp = Mr[((size_t)q * n + p) * m_pad + m], bqq = Mr[((size_t)q * n + q) * m_pad + m];
          const double h_pq = bpq + bqp, h_d = bpp - bqq;
          g00 += h_pq * h_pq; g01 += h_pq * h_d; g11 += h_d * h_d;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          g00 += __shfl_xor_sync(0xffffffffu, g00, o); g01 += __shfl_xor_sync(0xffffffffu, g01, o); g11 += __shfl_xor_sync(0xffffffffu, g11, o);
        }
        if (lane == 0) { sh[0][warp] = g00; sh[1][warp] = g01; sh[2][warp] = g11; }
        __syncthreads();
        if (tid == 0) {
          double a = 0, b = 0, d = 0;
          for (int w = 0; w < nw; ++w) { a += sh[0][w]; b += sh[1][w]; d += sh[2][w]; }
          const double diff = d - a;
          double angle = 0.0;
          if (!(fabs(b) < 1e-15 && fabs(diff) < 1e-15)) angle = 0.25 * atan2(2.0 * b, diff);  // jade.rs:174-179
          s_cs[0] = cos(angle); s_cs[1] = sin(angle); s_cs[2] = angle;
        }
        __syncthreads();
        const double cc = s_cs[0], ss = s_cs[1];
        max_theta = fmax(max_theta, fabs(s_cs[2]));
        // ---- V <- V G (jade.rs:188-197): v_p' = c v_p - s v_q ; v_q' = s v_p + c v_q
        if (tid < n) {
          const double vp = vsm[tid * n + p], vq = vsm[tid * n + q];
          vsm[tid * n + p] = cc * vp - ss * vq;
          vsm[tid * n + q] = ss * vp + cc * vq;
        }
        // ---- M' <- G^T M' G for every matrix: columns p, q then rows p, q (thread m <-> matrix m: coalesced)
        if (s_cs[2] != 0.0) {
          // batches of 8 independent element pairs: 16 loads in flight per thread before the first dependent store (a plain
          // load-compute-store loop serialises on L2 latency because the stores may alias the next loads)
          for (int m = tid; m < n_pairs; m += nt) {
            for (int k0 = 0; k0 < n; k0 += 8) {
              double xv[8], yv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                xv[u] = k < n ? Mr[((size_t)k * n + p) * m_pad + m] : 0.0;
                yv[u] = k < n ? Mr[((size_t)k * n + q) * m_pad + m] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k < n) {
                  Mr[((size_t)k * n + p) * m_pad + m] = cc * xv[u] - ss * yv[u];
                  Mr[((size_t)k * n + q) * m_pad + m] = ss * xv[u] + cc * yv[u];
                }
              }
            }
            for (int k0 = 0; k0 < n; k0 += 8) {
              double xv[8], yv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                xv[u] = k < n ? Mr[((size_t)p * n + k) * m_pad + m] : 0.0;
                yv[u] = k < n ? Mr[((size_t)q * n + k) * m_pad + m] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k < n) {
                  Mr[((size_t)p * n + k) * m_pad + m] = cc * xv[u] - ss * yv[u];
                  Mr[((size_t)q * n + k) * m_pad + m] = ss * xv[u] + cc * yv[u];
                }
              }
            }
          }
        }
        __syncthreads();
      }
    sweeps = iter + 1;
    if (max_theta < tol) break;  // uniform: every thread tracked the same angles
  }
  for (int e = tid; e < n * n; e += nt) V_out[e] = vsm[e];
  if (tid == 0) *sweeps_out = sweeps;
}

template <int NP>
int launch_cumulants(const double* d_x, int n, int64_t t_local, int64_t ld, int sm_count, cudaStream_t st, double* d_partial, int& n_tg,
                     int& pairs_padded) {
  using G = JadeGeom<NP>;
  static PerDeviceInt configured;
  configured.get([&] {
    PICARD_CUDA(cudaFuncSetAttribute(jade_cumulant_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    return 1;
  });
  const int n_pairs = n * (n + 1) / 2;
  const int n_pg = (n_pairs + G::PPC - 1) / G::PPC;
  const int64_t n_tiles = (t_local + G::BT - 1) / G::BT;
  int64_t tg = sm_count / n_pg;
  if (tg < 1) tg = 1;
  if (tg > n_tiles) tg = n_tiles;
  n_tg = (int)tg;
  pairs_padded = n_pg * G::PPC;
  if (!d_partial) return 0;  // sizing call
  CUtensorMap tmap = make_tmap(d_x, ld, t_local, n, NP);
  JadeParams p;
  p.n = n; p.n_pairs = n_pairs; p.t_local = t_local; p.n_tiles = n_tiles; p.partial = d_partial; p.n_pg = n_pg;
  jade_cumulant_kernel<NP><<<(unsigned)(n_tg * n_pg), G::NTHREADS, G::SMEM_BYTES, st>>>(tmap, p);
  PICARD_CUDA(cudaGetLastError());
  return 1;
}

int cumulants_dispatch(int np, const double* d_x, int n, int64_t t_local, int64_t ld, int sm_count, cudaStream_t st, double* d_partial, int& n_tg,
                       int& pairs_padded) {
  switch (np) {
    case 8: return launch_cumulants<8>(d_x, n, t_local, ld, sm_count, st, d_partial, n_tg, pairs_padded);
    case 16: return launch_cumulants<16>(d_x, n, t_local, ld, sm_count, st, d_partial, n_tg, pairs_padded);
    default: return launch_cumulants<32>(d_x, n, t_local, ld, sm_count, st, d_partial, n_tg, pairs_padded);
  }
}

// raw sums -> (allreduce) -> cumulant matrices in both layouts
void cumulants_device(const double* d_x, int n, int64_t t_local, int64_t ld, double t_total, picard_comm* comm, int sm_count, cudaStream_t st,
                      double* d_std, double* d_rot, int m_pad, picard_stats_t* stats) {
  if (n > 32)
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: the device JADE warm start supports at most 32 components (got " +
                                               std::to_string(n) + "); its cost grows like N^4 T");
  const int np = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
  const int n_pairs = n * (n + 1) / 2;
  int n_tg = 0, pairs_padded = 0;
  cumulants_dispatch(np, d_x, n, t_local, ld, sm_count, st, nullptr, n_tg, pairs_padded);
  DevBuf<double> partial((size_t)n_tg * pairs_padded * np * np), raw((size_t)n_pairs * n * n);
  stats->kernel_launches += cumulants_dispatch(np, d_x, n, t_local, ld, sm_count, st, partial.p, n_tg, pairs_padded);
  const int64_t total = (int64_t)n_pairs * n * n;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 4 * (int64_t)sm_count);
  jade_reduce_kernel<<<blocks, 256, 0, st>>>(partial.p, n_tg, pairs_padded, np, n, n_pairs, raw.p);
  PICARD_CUDA(cudaGetLastError());
  comm_allreduce_sum(comm, raw.p, (size_t)total, st);
  jade_finalize_kernel<<<blocks, 256, 0, st>>>(raw.p, n, n_pairs, m_pad, t_total, d_std, d_rot);
  PICARD_CUDA(cudaGetLastError());
  stats->kernel_launches += 2;
  PICARD_CUDA(cudaStreamSynchronize(st));  // partial / raw go out of scope
}

}  // namespace

void jade_cumulants_device(const double* d_x, int n, int64_t t_local, int64_t ld, double t_total, picard_comm* comm, int sm_count,
                           cudaStream_t st, double* out_host, picard_stats_t* stats) {
  const int n_pairs = n * (n + 1) / 2;
  DevBuf<double> d_std((size_t)n_pairs * n * n);
  cumulants_device(d_x, n, t_local, ld, t_total, comm, sm_count, st, d_std.p, nullptr, 0, stats);
  PICARD_CUDA(cudaMemcpyAsync(out_host, d_std.p, sizeof(double) * (size_t)n_pairs * n * n, cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaStreamSynchronize(st));
}

void jade_device(const double* d_x, int n, int64_t t_local, int64_t ld, double t_total, int64_t max_iter, double tol, bool verbose,
                 picard_comm* comm, int sm_count, cudaStream_t st, double* w_out, int64_t* sweeps_done, picard_stats_t* stats) {
  if (sweeps_done) *sweeps_done = 0;
  if (n < 2) {  // jade.rs:25-27
    for (int i = 0; i < n * n; ++i) w_out[i] = (i / (n > 0 ? n : 1) == i % (n > 0 ? n : 1)) ? 1.0 : 0.0;
    return;
  }
  const bool talk = verbose && (!comm || comm_rank(comm) == 0);
  const int n_pairs = n * (n + 1) / 2;
  const int m_pad = (n_pairs + 31) / 32 * 32;
  DevBuf<double> rot((size_t)n * n * m_pad), V((size_t)n * n), work(small::sym_decorrelation_work(n)), W((size_t)n * n);
  DevBuf<int> d_sweeps(1), d_status(1);
  rot.zero(st);
  cumulants_device(d_x, n, t_local, ld, t_total, comm, sm_count, st, nullptr, rot.p, m_pad, stats);
  if (talk) printf("JADE: %d cumulant matrices computed\n", n_pairs);  // jade.rs:32-34
  if (max_iter > 0) {
    jade_sweep_kernel<<<1, 1024, sizeof(double) * n * n, st>>>(rot.p, n, n_pairs, m_pad, (int)std::min<int64_t>(max_iter, 1 << 30), tol, V.p,
                                                                 d_sweeps.p);
    PICARD_CUDA(cudaGetLastError());
    stats->kernel_launches += 1;
  } else {
    stats->kernel_launches += small::set_identity(V.p, n, st);
    d_sweeps.zero(st);
  }
  stats->kernel_launches += small::sym_decorrelation(V.p, n, work.p, W.p, d_status.p, st);  // jade.rs:69
  int sweeps = 0, status = 0;
  PICARD_CUDA(cudaMemcpyAsync(&sweeps, d_sweeps.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaMemcpyAsync(&status, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaMemcpyAsync(w_out, W.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
  PICARD_CUDA(cudaStreamSynchronize(st));
  if (status != PICARD_OK) throw Error(PICARD_SINGULAR_MATRIX, "Singular matrix encountered during computation");
  if (sweeps_done) *sweeps_done = sweeps;
  if (talk && sweeps > 0 && sweeps < max_iter) printf("JADE converged after %d iterations\n", sweeps);
}

}  // namespace picard
