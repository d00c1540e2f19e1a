// N x N device kernels of the core loop.  See small.cuh.  Reference citations per kernel.
#include "small.cuh"
#include "loss_point.cuh"
#include "exact_div.h"

#include <cmath>

#include "pass.cuh"  // moment-buffer layout helpers

namespace picard {
namespace small {

namespace {

constexpr int BT1 = 1024;  // single-CTA kernels

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double r = lane < nw ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  double r = sh[32];
  __syncthreads();
  return r;
}
// fmax ignores NaN operands like Rust's f64::max in `fold(0.0, f64::max)` (core.rs:289, math.rs:42)
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double r = lane < nw ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  double r = sh[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double rust_signum(double v) {  // quirk Q8: +0 -> +1, -0 -> -1, NaN -> NaN
  if (v != v) return v;
  return signbit(v) ? -1.0 : 1.0;
}

// loss from reduced moments: core.rs:51-82 after the per-row sums
__device__ double loss_of_point(const CoreDims& d, const double* mom, const double* signs, bool* singular) {
  const int n = d.n;
  const double tf = d.t_total;
  double loss = 0.0;
  *singular = false;
  if (!d.ortho) {
    const double* ex = mom + mom_size(n);
    if (ex[1] == 0.0) { *singular = true; return 1e15; }
    loss = -ex[0];
  }
  const double* L = mom + mom_off_ll(n);
  const double* Sq = mom + mom_off_sq(n);
  for (int i = 0; i < n; ++i) {
    const double s = signs ? signs[i] : 1.0;
    loss += s * L[i] / tf;
    if (d.extended && !d.ortho) loss += 0.5 * Sq[i] / tf;
  }
  return loss;
}

// ---------------------------------------------------------------------------------------------------
// iteration front + L-BFGS update + direction, single CTA
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT1) front_kernel(FrontArgs a) {
  __shared__ double sh[33];
  __shared__ int sh_flag;
  const int n = a.d.n, nn = n * n, tid = threadIdx.x, nt = blockDim.x;
  const double tf = a.d.t_total;
  const bool ortho = a.d.ortho != 0, ext = a.d.extended != 0;
  const double* Gr = a.mom + mom_off_gr(n);
  const double* Hr = a.mom + mom_off_hr(n);
  const double* Sd = a.mom + mom_off_sd(n);
  const double* Sq = a.mom + mom_off_sq(n);

  const int m = a.d.m;
  int len = 0, head = 0;
  if (a.do_lbfgs != 2) {
  if (tid == 0) sh_flag = 0;
  __syncthreads();
  // ---- extended: kurtosis-sign estimate (core.rs:225-237)
  if (ext) {
    int change = 0;
    for (int i = tid; i < n; i += nt) {
      const double pm = Sd[i] / tf, gii = Gr[i * n + i] / tf;
      const double s = rust_signum(pm * a.C[i * n + i] - gii);
      if (!a.first_iter && s != a.old_signs[i]) change = 1;
      a.signs[i] = s;
      a.old_signs[i] = s;
    }
    if (change) atomicOr(&sh_flag, 1);
  } else {
    for (int i = tid; i < n; i += nt) a.signs[i] = 1.0;
  }
  __syncthreads();
  const int sign_change = sh_flag;
  // ---- g = Gr / T, sign-scaled rows, + C when not ortho (core.rs:218, 240-252)
  for (int e = tid; e < nn; e += nt) {
    const int i = e / n;
    double v = Gr[e] / tf;
    if (ext) {
      v *= a.signs[i];
      if (!ortho) v = v + a.C[e];
    }
    a.Gtmp[e] = v;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) a.hoff[i] = ortho ? a.Gtmp[i * n + i] : 1.0;  // core.rs:256-260
  __syncthreads();
  // ---- Hessian approximation (core.rs:263-277)
  if (ortho) {
    for (int e = tid; e < nn; e += nt) {
      const int i = e / n, j = e % n;
      const double pmi = (ext ? a.signs[i] : 1.0) * (Sd[i] / tf), pmj = (ext ? a.signs[j] : 1.0) * (Sd[j] / tf);
      a.H[e] = fmax(0.5 * (pmi + pmj - a.hoff[i] - a.hoff[j]), a.d.lambda_min);
    }
  } else {
    for (int e = tid; e < nn; e += nt) {
      const int i = e / n, j = e % n;
      a.H[e] = ext ? (a.signs[i] * Hr[e] + Sq[j]) / tf : Hr[e] / tf;
    }
    __syncthreads();
    // regularize_hessian (lbfgs.rs:155-171): sequential in-place semantics = per unordered pair, (i,j) then (j,i)
    const int npairs = n * (n - 1) / 2;
    for (int pidx = tid; pidx < npairs; pidx += nt) {
      // unrank (i < j)
      int i = (int)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * pidx)) * 0.5);
      while ((long long)i * (2 * n - i - 1) / 2 > pidx) --i;
      while ((long long)(i + 1) * (2 * n - i - 2) / 2 <= pidx) ++i;
      const int j = pidx - (int)((long long)i * (2 * n - i - 1) / 2) + i + 1;
      double hij = a.H[i * n + j], hji = a.H[j * n + i];
      const double four = 4.0 * a.hoff[i] * a.hoff[j];
      {
        const double diff = hij - hji, discr = sqrt(diff * diff + four), ev = 0.5 * (hij + hji - discr);
        if (ev < a.d.lambda_min) hij += a.d.lambda_min - ev;
      }
      {
        const double diff = hji - hij, discr = sqrt(diff * diff + four), ev = 0.5 * (hji + hij - discr);
        if (ev < a.d.lambda_min) hji += a.d.lambda_min - ev;
      }
      a.H[i * n + j] = hij;
      a.H[j * n + i] = hji;
    }
  }
  // ---- projection (core.rs:280-286) and norm (core.rs:289)
  double mx = 0.0;
  for (int e = tid; e < nn; e += nt) {
    const int i = e / n, j = e % n;
    double v;
    if (ortho) v = (a.Gtmp[e] - a.Gtmp[j * n + i]) / 2.0;
    else v = (i == j) ? a.Gtmp[e] - 1.0 : a.Gtmp[e];
    a.G[e] = v;
    mx = fmax(mx, fabs(v));
  }
  const double gnorm = block_max(mx, sh);
  if (tid == 0) {
    a.sc->gradient_norm = gnorm;
    a.sc->sign_change = sign_change;
  }
  if (!a.do_lbfgs) return;
  __syncthreads();

  // ---- L-BFGS memory update (core.rs:296-314, quirk Q9)
  len = a.sc->mem_len; head = a.sc->mem_head;
  const int have_prev = a.sc->have_prev_step, have_old = a.sc->have_g_old;
  __syncthreads();
  if (!a.first_iter && have_prev && have_old) {
    double part = 0.0;
    for (int e = tid; e < nn; e += nt) {
      const double yd = a.G[e] - a.G_old[e];
      a.q[e] = yd;
      part += a.S_prev[e] * yd;
    }
    const double dot = block_sum(part, sh);
    const double r = 1.0 / dot;
    if (isfinite(r)) {
      int slot;
      if (len < m) { slot = (head + len) % m; ++len; }
      else { slot = head; head = (head + 1) % m; }
      for (int e = tid; e < nn; e += nt) {
        a.mem_s[(size_t)slot * nn + e] = a.S_prev[e];
        a.mem_y[(size_t)slot * nn + e] = a.q[e];
      }
      if (tid == 0) a.mem_r[slot] = r;
    }
    if (tid == 0) { a.sc->have_prev_step = 0; a.sc->last_r = r; }
  }
  for (int e = tid; e < nn; e += nt) a.G_old[e] = a.G[e];
  // ---- sign change: loss with the new signs, flush memory (core.rs:317-331, quirk Q11)
  if (ext && sign_change) {
    if (tid == 0) {
      bool sing;
      const double l = loss_of_point(a.d, a.mom, a.signs, &sing);
      a.sc->current_loss = l;  // singular -> 1e15 and continue
    }
    len = 0; head = 0;
  }
  if (tid == 0) { a.sc->mem_len = len; a.sc->mem_head = head; a.sc->have_g_old = 1; }
  __syncthreads();

  } else {
    len = a.sc->mem_len; head = a.sc->mem_head;
    __syncthreads();
  }
  // ---- direction: two-loop recursion (lbfgs.rs:84-133)
  double* alist = a.mem_r + m;
  for (int e = tid; e < nn; e += nt) a.q[e] = a.G[e];
  __syncthreads();
  for (int k = len - 1; k >= 0; --k) {
    const int slot = (head + k) % m;
    const double* s = a.mem_s + (size_t)slot * nn;
    const double* y = a.mem_y + (size_t)slot * nn;
    double part = 0.0;
    for (int e = tid; e < nn; e += nt) part += s[e] * a.q[e];
    const double al = a.mem_r[slot] * block_sum(part, sh);
    if (tid == 0) alist[k] = al;
    for (int e = tid; e < nn; e += nt) a.q[e] = a.q[e] - al * y[e];
    __syncthreads();
  }
  // preconditioner
  if (ortho) {
    for (int e = tid; e < nn; e += nt) a.Gtmp[e] = a.q[e] / a.H[e];
    __syncthreads();
    for (int e = tid; e < nn; e += nt) {
      const int i = e / n, j = e % n;
      a.D[e] = (a.Gtmp[e] - a.Gtmp[j * n + i]) / 2.0;
    }
  } else {  // solve_hessian_system (lbfgs.rs:136-150, quirk Q15)
    for (int e = tid; e < nn; e += nt) {
      const int i = e / n, j = e % n, et = j * n + i;
      const double det = a.H[e] * a.H[et] - a.hoff[i] * a.hoff[j];
      double v = 0.0;
      if (fabs(det) > 1e-15) v = (a.H[et] * a.q[e] - a.hoff[i] * a.q[et]) / det;
      a.D[e] = v;
    }
  }
  __syncthreads();
  for (int k = 0; k < len; ++k) {
    const int slot = (head + k) % m;
    const double* s = a.mem_s + (size_t)slot * nn;
    const double* y = a.mem_y + (size_t)slot * nn;
    double part = 0.0;
    for (int e = tid; e < nn; e += nt) part += y[e] * a.D[e];
    const double beta = a.mem_r[slot] * block_sum(part, sh);
    const double cf = alist[k] - beta;
    for (int e = tid; e < nn; e += nt) a.D[e] = a.D[e] + cf * s[e];
    __syncthreads();
  }
  double dm = 0.0;
  for (int e = tid; e < nn; e += nt) {
    const double v = -a.D[e];
    a.D[e] = v;
    dm = fmax(dm, fabs(v));
  }
  const double nd = block_max(dm, sh);
  if (tid == 0) {
    a.sc->norm_d = nd;
    publish_scalars(a.sc, a.sc_map, a.seq);
  }
}

// ---------------------------------------------------------------------------------------------------
// The same iteration front on ONE THREAD-BLOCK CLUSTER of 16 CTAs (n <= 128, m <= 7): every matrix element is owned by one
// thread for the whole kernel (CTA = 8 rows, warp = one row, lane = 4 consecutive columns), so G, H, q, D and the whole L-BFGS
// history live in registers; what the single-CTA kernel above re-reads from L2 in ~45 dependent sweeps (100 us at n = 128) is
// loaded once.  Values at the transposed position that are plain functions of the raw moments (G^T, H^T, hoff_j, sign_j) are
// recomputed from the moments instead of exchanged; the one genuine exchange (q^T for the preconditioner) goes through L2
// behind a cluster barrier.  Dot products and maxima: warp shuffles, 8 per-warp partials in shared memory, the CTA's partial
// pushed into every CTA's slot array over distributed shared memory, one cluster barrier, a fixed-order sum of the 16 slots
// (identical on every CTA).  Same formulas and the same per-element operation order as front_kernel; only the order of the
// reductions differs (last-bit differences in the dot products).
// ---------------------------------------------------------------------------------------------------
constexpr int FC_CTAS = 16, FC_THREADS = 256, FC_M = 7;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_store(double* local_addr, uint32_t target_rank, double v) {
  uint32_t la = (uint32_t)__cvta_generic_to_shared(local_addr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(target_rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
struct ClusterRed {
  double slot[2][2][FC_CTAS];  // [buffer][value][source CTA]
  double warp[2][8];           // [value][warp] partials of this CTA
};
// Reduces two values over the whole cluster: value 0 with + (SUM0) or fmax, value 1 always with fmax.  Every thread returns the
// same bits.  One cluster barrier; the slot buffers alternate so that no barrier is needed after the final reads.
template <bool SUM0>
__device__ __forceinline__ void cluster_reduce2(double& v0, double& v1, ClusterRed* cr, int& buf, uint32_t rank) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double p0 = __shfl_xor_sync(0xffffffffu, v0, o), p1 = __shfl_xor_sync(0xffffffffu, v1, o);
    v0 = SUM0 ? v0 + p0 : fmax(v0, p0);
    v1 = fmax(v1, p1);
  }
  if (lane == 0) { cr->warp[0][warp] = v0; cr->warp[1][warp] = v1; }
  __syncthreads();
  if (warp == 0 && lane < FC_CTAS) {
    double c0 = cr->warp[0][0], c1 = cr->warp[1][0];
#pragma unroll
    for (int w = 1; w < FC_THREADS / 32; ++w) { c0 = SUM0 ? c0 + cr->warp[0][w] : fmax(c0, cr->warp[0][w]); c1 = fmax(c1, cr->warp[1][w]); }
    dsmem_store(&cr->slot[buf][0][rank], (uint32_t)lane, c0);
    dsmem_store(&cr->slot[buf][1][rank], (uint32_t)lane, c1);
  }
  cluster_sync_all();
  double r0 = cr->slot[buf][0][0], r1 = cr->slot[buf][1][0];
#pragma unroll
  for (int c = 1; c < FC_CTAS; ++c) { r0 = SUM0 ? r0 + cr->slot[buf][0][c] : fmax(r0, cr->slot[buf][0][c]); r1 = fmax(r1, cr->slot[buf][1][c]); }
  v0 = r0; v1 = r1;
  buf ^= 1;
}

__global__ void __cluster_dims__(FC_CTAS, 1, 1) __launch_bounds__(FC_THREADS, 1) front_cluster_kernel(FrontArgs a) {
  __shared__ ClusterRed cr;
  const int n = a.d.n, nn = n * n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = a.d.m;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0 && tid == 0;
  const double tf = a.d.t_total, lam = a.d.lambda_min;
  const bool ortho = a.d.ortho != 0, ext = a.d.extended != 0;
  const double* Gr = a.mom + mom_off_gr(n);
  const double* Hr = a.mom + mom_off_hr(n);
  const double* Sd = a.mom + mom_off_sd(n);
  const double* Sq = a.mom + mom_off_sq(n);
  const int i = 8 * (int)rank + warp, j0 = 4 * lane;  // this thread: row i, columns j0 .. j0 + 3
  const bool rowok = i < n;
  bool ok[4];
  int e[4], et[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    ok[u] = rowok && (j0 + u) < n;
    e[u] = ok[u] ? i * n + j0 + u : 0;
    et[u] = ok[u] ? (j0 + u) * n + i : 0;
  }
  int rbuf = 0;
  // ring state before this launch (read by every thread BEFORE the first cluster barrier; the leader rewrites it after it)
  int len = a.sc->mem_len, head = a.sc->mem_head;
  const int have_prev = a.sc->have_prev_step, have_old = a.sc->have_g_old;

  double G[4], H[4], HT[4], hoff_i = 1.0, hoff_j[4];
  int new_slot = -1;
  double new_r = 0.0;
#pragma unroll
  for (int u = 0; u < 4; ++u) { G[u] = 0.0; H[u] = 1.0; HT[u] = 1.0; hoff_j[u] = 1.0; }

  if (a.do_lbfgs != 2) {
    // ---- signs of the own row and of the own columns (core.rs:225-237), straight from the moments
    double sgn_i = 1.0, sgn_j[4] = {1.0, 1.0, 1.0, 1.0}, pm_i = 0.0, pm_j[4] = {0.0, 0.0, 0.0, 0.0};
    double change = 0.0;
    if (rowok) {
      pm_i = Sd[i] / tf;
      if (ext) {
        const double gii = Gr[i * n + i] / tf;
        sgn_i = rust_signum(pm_i * a.C[i * n + i] - gii);
        if (!a.first_iter && sgn_i != a.old_signs[i]) change = 1.0;
      }
    }
    __syncwarp();  // every lane has read old_signs[i] before lane 0 overwrites it
    if (rowok && lane == 0) {
      a.signs[i] = sgn_i;
      if (ext) a.old_signs[i] = sgn_i;
    }
    double gt[4] = {0, 0, 0, 0}, gtT[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int j = j0 + u;
      pm_j[u] = Sd[j] / tf;
      const double gjj_raw = Gr[j * n + j] / tf;
      if (ext) sgn_j[u] = rust_signum(pm_j[u] * a.C[j * n + j] - gjj_raw);
      // g = Gr / T, sign-scaled rows, + C when not ortho (core.rs:218, 240-252)
      double v = Gr[e[u]] / tf;
      if (ext) { v *= sgn_i; if (!ortho) v = v + a.C[e[u]]; }
      gt[u] = v;
      a.Gtmp[e[u]] = v;
      double vjj = gjj_raw;
      if (ext) { vjj *= sgn_j[u]; if (!ortho) vjj = vjj + a.C[j * n + j]; }
      hoff_j[u] = ortho ? vjj : 1.0;                                   // core.rs:256-260
      if (ortho) { double vt = Gr[et[u]] / tf; if (ext) vt *= sgn_j[u]; gtT[u] = vt; }
    }
    if (rowok) {
      double vii = Gr[i * n + i] / tf;
      if (ext) { vii *= sgn_i; if (!ortho) vii = vii + a.C[i * n + i]; }
      hoff_i = ortho ? vii : 1.0;
      if (lane == 0) a.hoff[i] = hoff_i;
    }
    // ---- Hessian approximation (core.rs:263-277) and regularize_hessian (lbfgs.rs:155-171: per unordered pair, (lo,hi) then (hi,lo))
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int j = j0 + u;
      if (ortho) {
        const double pmi = (ext ? sgn_i : 1.0) * pm_i, pmj = (ext ? sgn_j[u] : 1.0) * pm_j[u];
        H[u] = fmax(0.5 * (pmi + pmj - hoff_i - hoff_j[u]), lam);
        HT[u] = H[u];
      } else {
        const double he = ext ? (sgn_i * Hr[e[u]] + Sq[j]) / tf : Hr[e[u]] / tf;
        const double ht = ext ? (sgn_j[u] * Hr[et[u]] + Sq[i]) / tf : Hr[et[u]] / tf;
        if (i == j) { H[u] = he; HT[u] = he; }
        else {
          double hij = i < j ? he : ht, hji = i < j ? ht : he;  // (lo,hi) and (hi,lo)
          const double four = i < j ? 4.0 * hoff_i * hoff_j[u] : 4.0 * hoff_j[u] * hoff_i;
          {
            const double diff = hij - hji, discr = sqrt(diff * diff + four), ev = 0.5 * (hij + hji - discr);
            if (ev < lam) hij += lam - ev;
          }
          {
            const double diff = hji - hij, discr = sqrt(diff * diff + four), ev = 0.5 * (hji + hij - discr);
            if (ev < lam) hji += lam - ev;
          }
          H[u] = i < j ? hij : hji;
          HT[u] = i < j ? hji : hij;
        }
      }
      a.H[e[u]] = H[u];
    }
    // ---- projection (core.rs:280-286) and norm (core.rs:289)
    double mx = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int j = j0 + u;
      double v;
      if (ortho) v = (gt[u] - gtT[u]) / 2.0;
      else v = (i == j) ? gt[u] - 1.0 : gt[u];
      G[u] = v;
      a.G[e[u]] = v;
      mx = fmax(mx, fabs(v));
    }
    cluster_reduce2<false>(mx, change, &cr, rbuf, rank);
    const double gnorm = mx;
    const int sign_change = change != 0.0 ? 1 : 0;
    if (leader) { a.sc->gradient_norm = gnorm; a.sc->sign_change = sign_change; }
    if (!a.do_lbfgs) return;

    // ---- L-BFGS memory update (core.rs:296-314, quirk Q9)
    if (!a.first_iter && have_prev && have_old) {
      double yd[4], sp[4], part = 0.0, zero = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        yd[u] = 0.0; sp[u] = 0.0;
        if (!ok[u]) continue;
        yd[u] = G[u] - a.G_old[e[u]];
        sp[u] = a.S_prev[e[u]];
        part += sp[u] * yd[u];
      }
      cluster_reduce2<true>(part, zero, &cr, rbuf, rank);
      const double r = 1.0 / part;
      if (isfinite(r)) {
        int slot;
        if (len < m) { slot = (head + len) % m; ++len; }
        else { slot = head; head = (head + 1) % m; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) { a.mem_s[(size_t)slot * nn + e[u]] = sp[u]; a.mem_y[(size_t)slot * nn + e[u]] = yd[u]; }
        if (leader) a.mem_r[slot] = r;
        new_slot = slot; new_r = r;
      }
      if (leader) { a.sc->have_prev_step = 0; a.sc->last_r = r; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) a.G_old[e[u]] = G[u];
    // ---- sign change: loss with the new signs, flush memory (core.rs:317-331, quirk Q11)
    if (ext && sign_change) {
      if (leader) {
        // the leader's own row-0 sign is in place; the other rows' signs were written by other CTAs before the barrier of the
        // first reduction, which every CTA has passed
        bool sing;
        const double l = loss_of_point(a.d, a.mom, a.signs, &sing);
        a.sc->current_loss = l;  // singular -> 1e15 and continue
      }
      len = 0; head = 0;
    }
    if (leader) { a.sc->mem_len = len; a.sc->mem_head = head; a.sc->have_g_old = 1; }
  } else {
    // direction only (test hook): G, H, hoff are in place
    if (rowok) hoff_i = a.hoff[i];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) { G[u] = a.G[e[u]]; H[u] = a.H[e[u]]; HT[u] = a.H[et[u]]; hoff_j[u] = a.hoff[j0 + u]; }
  }

  // ---- direction: two-loop recursion (lbfgs.rs:84-133) with the whole history in registers
  double sk[FC_M][4], yk[FC_M][4], rk[FC_M], alist[FC_M];
#pragma unroll
  for (int k = 0; k < FC_M; ++k) {
    rk[k] = 0.0; alist[k] = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { sk[k][u] = 0.0; yk[k][u] = 0.0; }
    if (k < len) {
      const int slot = (head + k) % m;
      rk[k] = slot == new_slot ? new_r : a.mem_r[slot];  // the pair pushed by this launch: r is known to every thread, and each thread
#pragma unroll                                            // reads back only the elements it wrote itself
      for (int u = 0; u < 4; ++u)
        if (ok[u]) { sk[k][u] = a.mem_s[(size_t)slot * nn + e[u]]; yk[k][u] = a.mem_y[(size_t)slot * nn + e[u]]; }
    }
  }
  double q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) q[u] = G[u];
#pragma unroll
  for (int k = FC_M - 1; k >= 0; --k) {
    if (k >= len) continue;  // uniform over the cluster
    double part = 0.0, zero = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) part += sk[k][u] * q[u];
    cluster_reduce2<true>(part, zero, &cr, rbuf, rank);
    const double al = rk[k] * part;
    alist[k] = al;
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = q[u] - al * yk[k][u];
  }
  // preconditioner: the value at the transposed position belongs to another CTA -> one exchange through L2
  double D[4];
  if (ortho) {
    double gq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { gq[u] = ok[u] ? q[u] / H[u] : 0.0; if (ok[u]) a.Gtmp[e[u]] = gq[u]; }
    __threadfence();
    cluster_sync_all();
#pragma unroll
    for (int u = 0; u < 4; ++u) D[u] = ok[u] ? (gq[u] - __ldcg(a.Gtmp + et[u])) / 2.0 : 0.0;
  } else {  // solve_hessian_system (lbfgs.rs:136-150, quirk Q15)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) a.q[e[u]] = q[u];
    __threadfence();
    cluster_sync_all();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      D[u] = 0.0;
      if (!ok[u]) continue;
      const double qt = __ldcg(a.q + et[u]);
      const double det = H[u] * HT[u] - hoff_i * hoff_j[u];
      if (fabs(det) > 1e-15) D[u] = (HT[u] * q[u] - hoff_i * qt) / det;
    }
  }
#pragma unroll
  for (int k = 0; k < FC_M; ++k) {
    if (k >= len) continue;
    double part = 0.0, zero = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) part += yk[k][u] * D[u];
    cluster_reduce2<true>(part, zero, &cr, rbuf, rank);
    const double beta = rk[k] * part;
    const double cf = alist[k] - beta;
#pragma unroll
    for (int u = 0; u < 4; ++u) D[u] = D[u] + cf * sk[k][u];
  }
  double dm = 0.0, zero = 0.0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const double v = -D[u];
    if (ok[u]) { a.D[e[u]] = v; dm = fmax(dm, fabs(v)); }
  }
  cluster_reduce2<false>(dm, zero, &cr, rbuf, rank);
  if (leader) {
    a.sc->norm_d = dm;
    publish_scalars(a.sc, a.sc_map, a.seq);
  }
}

__global__ void loss_kernel(CoreDims d, const double* mom, const double* signs, CoreScalars* sc, int which, CoreScalars* sc_map,
                            unsigned long long seq) {
  bool sing;
  const double l = loss_of_point_warp(d, mom, signs, &sing);
  if (threadIdx.x != 0) return;
  if (which == 0) {
    sc->new_loss = l;
    sc->accept = (l < sc->current_loss) ? 1 : 0;
  } else {
    sc->current_loss = l;
    sc->loss_singular = sing ? 1 : 0;
  }
  publish_scalars(sc, sc_map, seq);
}

__global__ void accept_kernel(const double* D, double alpha, double* S_prev, int64_t count, CoreScalars* sc) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
    S_prev[e] = D[e] * alpha;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->current_loss = sc->new_loss;
    sc->have_prev_step = 1;
  }
}
__global__ void negate_kernel(const double* G, double* D, int64_t count, CoreScalars* sc) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) D[e] = -G[e];
  if (blockIdx.x == 0 && threadIdx.x == 0) { sc->mem_len = 0; sc->mem_head = 0; sc->norm_d = sc->gradient_norm; }
}
__global__ void scale_kernel(const double* A, double* B, int64_t count, double alpha) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) B[e] = A[e] * alpha;
}
__global__ void eye_plus_kernel(const double* D, double alpha, double* M, int n) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x)
    M[e] = ((e / n == e % n) ? 1.0 : 0.0) + alpha * D[e];
}
__global__ void identity_kernel(double* A, int n) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) A[e] = (e / n == e % n) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// small dense matmul: 32 x 32 output tile per CTA of 256 threads (2 x 2 per thread), k in chunks of 32
// ---------------------------------------------------------------------------------------------------
template <bool TRANS_B>
__device__ __forceinline__ void tile_mm(const double* A, int lda, const double* B, int ldb, int m, int k, int n, int bi, int bj,
                                        double acc[2][2]) {
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
  for (int k0 = 0; k0 < k; k0 += 32) {
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      const int gi = bi * 32 + r, gk = k0 + c;
      As[r][c] = (gi < m && gk < k) ? A[(size_t)gi * lda + gk] : 0.0;
      const int gkk = k0 + r, gj = bj * 32 + c;
      double bv = 0.0;
      if (gkk < k && gj < n) bv = TRANS_B ? B[(size_t)gj * ldb + gkk] : B[(size_t)gkk * ldb + gj];
      Bs[r][c] = bv;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double a0 = As[ty][kk], a1 = As[ty + 16][kk], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
}

template <bool TRANS_B>
__global__ void __launch_bounds__(256) matmul_kernel(const double* A, const double* B, double* C, int m, int k, int n, double alpha,
                                                     int add_identity) {
  double acc[2][2];
  tile_mm<TRANS_B>(A, k, B, TRANS_B ? k : n, m, k, n, blockIdx.y, blockIdx.x, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = blockIdx.y * 32 + ty + 16 * a, j = blockIdx.x * 32 + tx + 16 * b;
      if (i < m && j < n) {
        double v = alpha * acc[a][b];
        if (add_identity && i == j) v += 1.0;
        C[(size_t)i * n + j] = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------------
// matrix_exp (math.rs:38-74) [+ W' = expm(alpha D) W, core.rs:125] as ONE kernel, ROW-partitioned: term_k = term_{k-1} A_s / k
// only needs the same rows of term_{k-1}, so each CTA owns 8 rows of term / result (shared memory) and reads A_s
// (= D alpha 2^-s, formed on the fly from D) through L2; products on DMMA.  The only cross-CTA dependency of the Taylor
// loop is the reference's early exit on the GLOBAL max |term_k| < 1e-16: every CTA publishes its local max per term and
// reads the others' (flags in global memory) -- no grid-wide barrier per term (grid.sync() cost ~10 us x ~13 terms in the
// previous version).  Squarings (s > 0, rare) exchange the result through global memory with one barrier each.
// Cooperative launch only for the co-residency guarantee of the flag waits.
// ---------------------------------------------------------------------------------------------------
constexpr int EXPM_ROWS = 8;
constexpr int EXPM_MAX_CTAS = 32;

// term / k exactly as the reference's division (math.rs:60) without the generic DDIV sequence (4 per thread and term: it was a
// third of the Taylor loop's instructions): div_by_count() of exact_div.h, checked on the host by tests/test_exact_div_host.py.
// acc[t][e] (t < 4) += T (8 x n, shared, pitch P) * B (n x n, global; element transform (b * mul0) * mul1) for column blocks
// cb = warp + 8 t; thread (j, c): rows c, columns 8 cb + 2 j + e.  B may have been written by other CTAs: plain coherent loads.
// ldb: leading dimension of B (n for global matrices, the padded pitch for the shared-memory copy of A_s).
template <bool CG>
__device__ __forceinline__ void rows_times_general(const double* T, int P, const double* B, int ldb, int n, double mul0, double mul1, double acc[4][2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j = lane & 3, c = lane >> 2;
  const int ncb = (n + 7) >> 3;
#pragma unroll
  for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 8
  for (int k0 = 0; k0 < n; k0 += 4) {
    const bool kok = k0 + j < n;
    const double a = kok ? T[c * P + k0 + j] : 0.0;
    double b[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int col = 8 * (warp + 8 * t) + c;
      b[t] = (kok && warp + 8 * t < ncb && col < n) ? ((CG ? __ldcg(B + (size_t)(k0 + j) * ldb + col) : B[(size_t)(k0 + j) * ldb + col]) * mul0) * mul1 : 0.0;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (warp + 8 * t < ncb)  // warp-uniform
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a), "d"(b[t]));
  }
}

// The same product when n is a multiple of 8 and B needs no scaling (A_s copied to shared memory, or W): no bounds logic, no
// per-element address arithmetic -- the general loop above spends ~40 integer instructions per DMMA, which is what bound the
// transform kernels (ncu: DMMA pipe 16 % busy).  NV = number of valid column blocks of this warp (warp-uniform).  Same DMMA
// sequence per accumulator as the general loop: bit-identical results.
template <int NV, bool CG>
__device__ __forceinline__ void rows_times_dense(const double* __restrict__ T, int P, const double* B, int ldb, int n, double acc[4][2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j = lane & 3, c = lane >> 2;
  const double* ap = T + c * P + j;
  const double* bp = B + (size_t)j * ldb + 8 * warp + c;
  const size_t kstep = (size_t)4 * ldb;
#pragma unroll
  for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 8
  for (int k0 = 0; k0 < n; k0 += 4) {
    const double a = ap[k0];
    double b[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) b[t] = CG ? __ldcg(bp + 64 * t) : bp[64 * t];
    bp += kstep;
#pragma unroll
    for (int t = 0; t < NV; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a), "d"(b[t]));
  }
}
// CG: B was written by other CTAs of this launch (the squaring phase): read it around L1 (ld.global.cg), which is not coherent
// across SMs -- a line of the same buffer cached two squarings ago must not be served again.
template <bool CG = false>
__device__ __forceinline__ void rows_times(const double* T, int P, const double* B, int ldb, int n, double mul0, double mul1, double acc[4][2]) {
  if ((n & 7) == 0 && mul0 == 1.0 && mul1 == 1.0) {  // uniform over the grid
    const int warp = threadIdx.x >> 5, ncb = n >> 3;
    const int nv = warp < ncb ? (ncb - warp + 7) >> 3 : 0;  // column blocks warp, warp + 8, ... < ncb
    switch (nv) {
      case 1: rows_times_dense<1, CG>(T, P, B, ldb, n, acc); break;
      case 2: rows_times_dense<2, CG>(T, P, B, ldb, n, acc); break;
      case 3: rows_times_dense<3, CG>(T, P, B, ldb, n, acc); break;
      case 4: rows_times_dense<4, CG>(T, P, B, ldb, n, acc); break;
      default:
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    }
    return;
  }
  rows_times_general<CG>(T, P, B, ldb, n, mul0, mul1, acc);
}

__global__ void __launch_bounds__(256) expm_rows_kernel(const double* __restrict__ D, double alpha, double inv_scale, double first_norm, int s,
                                                        int n, double* flags, unsigned int* bar, double* Rg0, double* Rg1, double* out,
                                                        const double* __restrict__ W, double* Wt, int as_in_smem) {
  extern __shared__ double esm[];
  const int P = n + 4, PA = n + 8;
  double* Tc = esm; double* Tn = esm + EXPM_ROWS * P; double* R = esm + 2 * EXPM_ROWS * P;
  double* Asm = esm + 3 * EXPM_ROWS * P;  // n x PA copy of A_s when it fits (n <= 128): B fragments at shared-memory latency
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = lane & 3, c = lane >> 2;
  const int G = gridDim.x, row0 = blockIdx.x * EXPM_ROWS, ncb = (n + 7) >> 3;
  // term_1 = A_s rows ; result = I + A_s   (A_s = (D alpha) / 2^s, math.rs:47-53)
  for (int e = tid; e < EXPM_ROWS * n; e += blockDim.x) {
    const int r = e / n, col = e % n, grow = row0 + r;
    const double v = grow < n ? (D[(size_t)grow * n + col] * alpha) * inv_scale : 0.0;
    Tc[r * P + col] = v;
    R[r * P + col] = ((grow == col) ? 1.0 : 0.0) + v;
  }
  if (as_in_smem)
    for (int r = warp; r < n; r += 8)  // (no integer division per element: it was 15 % of the kernel)
      for (int col = lane; col < n; col += 32) Asm[r * PA + col] = (D[(size_t)r * n + col] * alpha) * inv_scale;
  __syncthreads();
  // Taylor terms.  The reference stops after the first term whose GLOBAL max is below 1e-16 (math.rs:58-66).  Term k is
  // computed speculatively, then the published maxima of term k-1 (written one whole term ago: no waiting in practice)
  // decide whether it exists; a term that does not exist is discarded, so the result is exactly the reference's sum.
  if (!(first_norm < 1e-16)) {
    // one block barrier per term, as in expm_multi_kernel below: the maxima of term k-1 are requested before the product and
    // consumed after it, term k goes to Tn unconditionally, one reduction carries both maxima
    __shared__ double red[2][2][8];
    int rbuf = 0;
    for (int k = 2; k <= 30; ++k) {
      double gv = 0.0;
      const bool poll = k > 2 && tid < G;
      const volatile double* fprev = flags + (size_t)(k - 1) * G + tid;
      if (poll) gv = *fprev;
      double acc[4][2];
      if (as_in_smem) rows_times(Tc, P, Asm, PA, n, 1.0, 1.0, acc);  // a shared-memory pointer the compiler can see: LDS, not generic loads
      else rows_times(Tc, P, D, n, n, alpha, inv_scale, acc);
      if (poll)
        while (gv != gv) gv = *fprev;  // NaN = not published yet
      double mx = 0.0;
      const double kk = (double)k, rcp = 1.0 / kk;
      double val[4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * t) + 2 * j + e;
          val[t][e] = 0.0;
          if (warp + 8 * t < ncb && col < n) {
            const double v = div_by_count(acc[t][e], kk, rcp);
            val[t][e] = v;
            Tn[c * P + col] = v;
            if (row0 + c < n) mx = fmax(mx, fabs(v));
          }
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        gv = fmax(gv, __shfl_xor_sync(0xffffffffu, gv, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane == 0) { red[rbuf][0][warp] = gv; red[rbuf][1][warp] = mx; }
      __syncthreads();
      double gprev = 0.0, own = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { gprev = fmax(gprev, red[rbuf][0][w8]); own = fmax(own, red[rbuf][1][w8]); }
      rbuf ^= 1;
      if (k > 2 && gprev < 1e-16) break;  // uniform over the whole grid: every CTA reduces the same published values
      if (tid == 0) *reinterpret_cast<volatile double*>(flags + (size_t)k * G + blockIdx.x) = own;
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * t) + 2 * j + e;
          if (warp + 8 * t < ncb && col < n) R[c * P + col] += val[t][e];
        }
      double* t = Tc; Tc = Tn; Tn = t;
    }
    __syncthreads();  // R complete before the squaring / the output below reads other threads' entries
  }
  // squaring (math.rs:69-71): result <- result * result, s times; the full result goes through global memory
  double* Rg = Rg0; double* Rn = Rg1;
  for (int q = 0; q < s; ++q) {
    for (int e = tid; e < EXPM_ROWS * n; e += blockDim.x) {
      const int r = e / n, col = e % n;
      if (row0 + r < n) Rg[(size_t)(row0 + r) * n + col] = R[r * P + col];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicAdd(bar, 1u);
      while (*reinterpret_cast<volatile unsigned int*>(bar) < (unsigned)(G * (q + 1))) {}
      __threadfence();
    }
    __syncthreads();
    double acc[4][2];
    rows_times<true>(R, P, Rg, n, n, 1.0, 1.0, acc);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (warp + 8 * t < ncb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * t) + 2 * j + e;
          if (col < n) R[c * P + col] = acc[t][e];
        }
    __syncthreads();
    double* t = Rg; Rg = Rn; Rn = t;
  }
  if (out != nullptr)
    for (int e = tid; e < EXPM_ROWS * n; e += blockDim.x) {
      const int r = e / n, col = e % n;
      if (row0 + r < n) out[(size_t)(row0 + r) * n + col] = R[r * P + col];
    }
  if (W != nullptr && Wt != nullptr) {  // W' rows = result rows * W   (core.rs:125)
    double acc[4][2];
    rows_times(R, P, W, n, n, 1.0, 1.0, acc);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (warp + 8 * t < ncb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * t) + 2 * j + e;
          if (col < n && row0 + c < n) Wt[(size_t)(row0 + c) * n + col] = acc[t][e];
        }
  }
}

// All candidate transforms of one backtracking line search from ONE Taylor run (core.rs:118-125 for alpha = a0 / 2^t,
// t = 0 .. NC-1).  With s = 0 (max |a0 D| <= 1) the reference's series for candidate t is the a0 series with term k scaled
// by 2^{-k t} -- EXACTLY, powers of two commute with every rounding -- and its early exit is the first k with
// max |term_k| 2^{-k t} < 1e-16.  So one pass over the terms accumulates result_t = I + sum_k term_k 2^{-k t} for every t
// in the reference's order (bit-identical to running the series per try), then W'_t = result_t W.  NT = 8-column blocks
// per warp (2 for n <= 128, 4 for n <= 256).
template <int NT>
__global__ void __launch_bounds__(256) expm_multi_kernel(const double* __restrict__ D, double alpha, double first_norm, int n, int nc,
                                                         double* flags, const double* __restrict__ W, double* Wt_all, int as_in_smem) {
  extern __shared__ double esm[];
    const int P = n + 4, PA = n + 8;
  double* Tc = esm; double* Tn = esm + EXPM_ROWS * P;
  double* Asm = esm + 2 * EXPM_ROWS * P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = lane & 3, c = lane >> 2;
  const int G = gridDim.x, row0 = blockIdx.x * EXPM_ROWS, ncb = (n + 7) >> 3;
  for (int e = tid; e < EXPM_ROWS * n; e += blockDim.x) {
    const int r = e / n, col = e % n, grow = row0 + r;
    Tc[r * P + col] = grow < n ? D[(size_t)grow * n + col] * alpha : 0.0;
  }
  if (as_in_smem)
    for (int r = warp; r < n; r += 8)  // (no integer division per element: it was 15 % of the kernel)
      for (int col = lane; col < n; col += 32) Asm[r * PA + col] = D[(size_t)r * n + col] * alpha;
  __syncthreads();
  // per-thread result entries of every candidate: rows row0 + c, columns 8 (warp + 8 t') + 2 j + e
  double rt[EXPM_NC][NT][2];
  double pw[EXPM_NC];      // 2^{-k t} of the current k
  double step[EXPM_NC];    // 2^{-t}
  bool alive[EXPM_NC];
#pragma unroll
  for (int t = 0; t < EXPM_NC; ++t) {
    step[t] = 1.0 / (double)(1u << t);
    pw[t] = step[t];  // k = 1
    alive[t] = t < nc && !(first_norm * pw[t] < 1e-16);  // exit test after term 1
#pragma unroll
    for (int tt = 0; tt < NT; ++tt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * (warp + 8 * tt) + 2 * j + e;
        const double a1 = (col < n) ? Tc[c * P + col] : 0.0;
        rt[t][tt][e] = ((row0 + c == col) ? 1.0 : 0.0) + a1 * pw[t];
      }
  }
  // One block barrier per term: the published maxima of term k-1 are requested BEFORE the product of term k and consumed after it
  // (their round trip to L2 hides behind the DMMAs), term k goes to Tn unconditionally (a term that turns out not to exist is never
  // read), and ONE reduction carries both maxima -- the device-wide one of term k-1 that decides the reference's `break` per
  // candidate, and this CTA's own of term k that is published next.  Double-buffered reduction slots: no barrier after the reads.
  __shared__ double red[2][2][8];
  int rbuf = 0;
  for (int k = 2; k <= 30; ++k) {
    double gv = 0.0;
    const bool poll = k > 2 && tid < G;
    const volatile double* fprev = flags + (size_t)(k - 1) * G + tid;
    if (poll) gv = *fprev;
    double acc[4][2];
    if (as_in_smem) rows_times(Tc, P, Asm, PA, n, 1.0, 1.0, acc);  // a shared-memory pointer the compiler can see: LDS, not generic loads
    else rows_times(Tc, P, D, n, n, alpha, 1.0, acc);
    if (poll)
      while (gv != gv) gv = *fprev;  // NaN = not published yet (rare: it was written one whole term ago)
    double mx = 0.0;
    const double kk = (double)k, rcp = 1.0 / kk;
    double val[NT][2];
#pragma unroll
    for (int tt = 0; tt < NT; ++tt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * (warp + 8 * tt) + 2 * j + e;
        val[tt][e] = 0.0;
        if (warp + 8 * tt < ncb && col < n) {
          const double v = div_by_count(acc[tt][e], kk, rcp);
          val[tt][e] = v;
          Tn[c * P + col] = v;
          if (row0 + c < n) mx = fmax(mx, fabs(v));
        }
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      gv = fmax(gv, __shfl_xor_sync(0xffffffffu, gv, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red[rbuf][0][warp] = gv; red[rbuf][1][warp] = mx; }
    __syncthreads();  // also: Tn complete, every warp done with Tc
    double gprev = 0.0, own = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) { gprev = fmax(gprev, red[rbuf][0][w8]); own = fmax(own, red[rbuf][1][w8]); }
    rbuf ^= 1;
    if (k > 2) {  // the reference's `break` for each candidate, on the published global max of term k-1
#pragma unroll
      for (int t = 0; t < EXPM_NC; ++t)
        if (alive[t] && gprev * pw[t] < 1e-16) alive[t] = false;  // pw[t] still holds 2^{-(k-1) t}
    }
    bool any = false;
#pragma unroll
    for (int t = 0; t < EXPM_NC; ++t) { pw[t] *= step[t]; any = any || alive[t]; }  // now 2^{-k t}
    if (!any) break;  // uniform over the grid: every CTA reduced the same published values
    if (tid == 0) *reinterpret_cast<volatile double*>(flags + (size_t)k * G + blockIdx.x) = own;  // the value is the message: no fence
#pragma unroll
    for (int tt = 0; tt < NT; ++tt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * (warp + 8 * tt) + 2 * j + e;
        if (warp + 8 * tt < ncb && col < n) {
#pragma unroll
          for (int t = 0; t < EXPM_NC; ++t)
            if (alive[t]) rt[t][tt][e] += val[tt][e] * pw[t];
        }
      }
    double* tsw = Tc; Tc = Tn; Tn = tsw;
  }
  // W'_t rows = result_t rows * W; the two row buffers alternate, so one barrier per candidate is enough
  __syncthreads();
#pragma unroll
  for (int t = 0; t < EXPM_NC; ++t) {
    if (t >= nc) break;
    double* buf = (t & 1) ? Tn : Tc;
#pragma unroll
    for (int tt = 0; tt < NT; ++tt)
      if (warp + 8 * tt < ncb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * tt) + 2 * j + e;
          if (col < n) buf[c * P + col] = rt[t][tt][e];
        }
    __syncthreads();
    double acc[4][2];
    rows_times(buf, P, W, n, n, 1.0, 1.0, acc);
    double* dst = Wt_all + (size_t)t * n * n;
#pragma unroll
    for (int tt = 0; tt < NT; ++tt)
      if (warp + 8 * tt < ncb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * (warp + 8 * tt) + 2 * j + e;
          if (col < n && row0 + c < n) dst[(size_t)(row0 + c) * n + col] = acc[tt][e];
        }
  }
}

// ---------------------------------------------------------------------------------------------------
// signed log-determinant: LU with partial pivoting, single CTA, in `work` (n x n)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT1) lu_logdet_kernel(const double* A, int n, double* work, double* out2) {
  __shared__ double shv[32];
  __shared__ int shi[32];
  __shared__ int piv_row;
  __shared__ double piv_val;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int e = tid; e < n * n; e += nt) work[e] = A[e];
  __syncthreads();
  double logabs = 0.0, sign = 1.0;
  for (int k = 0; k < n; ++k) {
    double bv = -1.0; int bi = n;
    for (int i = k + tid; i < n; i += nt) {
      const double v = fabs(work[(size_t)i * n + k]);
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { shv[warp] = bv; shi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double v = shv[0]; int idx = shi[0];
      for (int w = 1; w < nw; ++w) if (shv[w] > v || (shv[w] == v && shi[w] < idx)) { v = shv[w]; idx = shi[w]; }
      piv_row = idx; piv_val = v;
    }
    __syncthreads();
    const int p = piv_row;
    if (!(piv_val > 0.0)) { sign = 0.0; break; }  // exactly singular (or NaN column)
    if (p != k) {
      for (int j = tid; j < n; j += nt) {
        const double t = work[(size_t)k * n + j];
        work[(size_t)k * n + j] = work[(size_t)p * n + j];
        work[(size_t)p * n + j] = t;
      }
      sign = -sign;
    }
    __syncthreads();
    const double d = work[(size_t)k * n + k];
    if (d < 0.0) sign = -sign;
    logabs += log(fabs(d));
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      work[(size_t)i * n + j] -= (work[(size_t)i * n + k] / d) * work[(size_t)k * n + j];
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (sign == 0.0) { out2[0] = -INFINITY; out2[1] = 0.0; }
    else { out2[0] = logabs; out2[1] = sign; }
  }
}

// ---------------------------------------------------------------------------------------------------
// cyclic Jacobi eigensolver, single CTA, round-robin parallel ordering
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT1) jacobi_kernel(double* A, int n, double* V, double* evals, double* tmp) {
  __shared__ double sh[33];
  __shared__ double cs_c[256], cs_s[256];
  __shared__ int pr_p[256], pr_q[256];
  __shared__ int order[512];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int np = n + (n & 1);  // players (one dummy if n is odd)
  const int half = np / 2;
  for (int e = tid; e < n * n; e += nt) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < n * n; e += nt) {
      const double v = A[e];
      if (e / n == e % n) dg += v * v; else off += v * v;
    }
    off = block_sum(off, sh);
    dg = block_sum(dg, sh);
    if (off <= 1e-32 * dg || off == 0.0) break;
    for (int r = 0; r < np - 1; ++r) {
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = np - 1; q = r; }
        else { p = (r + tid) % (np - 1); q = (r - tid + (np - 1)) % (np - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(size_t)p * n + q];
          if (apq != 0.0) {
            const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else { q = -1; }
        pr_p[tid] = p; pr_q[tid] = q; cs_c[tid] = c; cs_s[tid] = s;
      }
      __syncthreads();
      for (int e = tid; e < half * n; e += nt) {  // rows: A <- J^T A
        const int pr = e / n, jj = e % n, p = pr_p[pr], q = pr_q[pr];
        if (q < 0) continue;
        const double c = cs_c[pr], s = cs_s[pr];
        const double ap = A[(size_t)p * n + jj], aq = A[(size_t)q * n + jj];
        A[(size_t)p * n + jj] = c * ap - s * aq;
        A[(size_t)q * n + jj] = s * ap + c * aq;
      }
      __syncthreads();
      for (int e = tid; e < half * n; e += nt) {  // columns: A <- A J, V <- V J
        const int pr = e % half, i = e / half, p = pr_p[pr], q = pr_q[pr];
        if (q < 0) continue;
        const double c = cs_c[pr], s = cs_s[pr];
        const double ap = A[(size_t)i * n + p], aq = A[(size_t)i * n + q];
        A[(size_t)i * n + p] = c * ap - s * aq;
        A[(size_t)i * n + q] = s * ap + c * aq;
        const double vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q];
        V[(size_t)i * n + p] = c * vp - s * vq;
        V[(size_t)i * n + q] = s * vp + c * vq;
      }
      __syncthreads();
    }
  }
  // ascending order of the diagonal
  if (tid == 0) {
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i) {
      const int o = order[i]; const double v = A[(size_t)o * n + o];
      int k = i - 1;
      while (k >= 0 && A[(size_t)order[k] * n + order[k]] > v) { order[k + 1] = order[k]; --k; }
      order[k + 1] = o;
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) evals[i] = A[(size_t)order[i] * n + order[i]];
  for (int e = tid; e < n * n; e += nt) tmp[e] = V[(size_t)(e / n) * n + order[e % n]];
  __syncthreads();
  for (int e = tid; e < n * n; e += nt) V[e] = tmp[e];
}


// ---------------------------------------------------------------------------------------------------
// symmetric eigensolver for n > 32: Householder tridiagonalisation + implicit-shift QL (the tred2 / tql2 scheme),
// single CTA.  The cyclic Jacobi kernel above needs ~10 sweeps x (n - 1) rounds x 3 barriers with O(n^2) global
// traffic each (87 ms at n = 128, 0.7 s at n = 256 on B200); here the O(n^3) parts (rank-2 updates, accumulation of
// the reflectors, application of the plane rotations) are data-parallel over the CTA and only the O(n^2) scalar QL
// recurrence is sequential (one thread, rotation parameters handed over through shared memory).
// A (n x n, symmetric, destroyed), Z (n x n work), V out: eigenvectors in COLUMNS, evals ascending.
// ---------------------------------------------------------------------------------------------------
constexpr int EIG_MAX_N = 512;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(BT1) tridiag_ql_kernel(double* __restrict__ A, int n, double* __restrict__ Z, double* __restrict__ V,
                                                         double* __restrict__ evals, unsigned long long* __restrict__ stamps) {
  __shared__ double sh[33];
  __shared__ double d[EIG_MAX_N], e[EIG_MAX_N], v[EIG_MAX_N], q[EIG_MAX_N], cs[EIG_MAX_N], sn[EIG_MAX_N];
  __shared__ int rank_of[EIG_MAX_N];
  __shared__ double s_scal[4];
  __shared__ int s_int[4];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;

  if (stamps && threadIdx.x == 0) stamps[0] = gtimer();
  // ---- phase 1: A = Q T Q^T, reflector k annihilates A[k+2.., k]; v_k kept in column k (rows k+1..), beta_k in q? no: in cs[]
  for (int k = 0; k + 2 < n; ++k) {
    const int m = n - k - 1;  // trailing block A[k+1.., k+1..]
    double part = 0.0;
    for (int i = tid; i < m; i += nt) { const double x = A[(size_t)(k + 1 + i) * n + k]; v[i] = x; part += x * x; }
    const double sigma = block_sum(part, sh);
    const double x0 = v[0];
    const double tail = sigma - x0 * x0;
    double beta = 0.0, alpha = x0;
    if (tail > 0.0 && sigma > 0.0) {
      alpha = -copysign(sqrt(sigma), x0);
      beta = 1.0 / (sigma - alpha * x0);  // 2 / (v^T v) with v = x - alpha e1
      __syncthreads();
      if (tid == 0) v[0] = x0 - alpha;
    }
    __syncthreads();
    if (tid == 0) { d[k] = A[(size_t)k * n + k]; e[k] = alpha; cs[k] = beta; }
    if (beta != 0.0) {
      // p = beta * A22 v : one warp per row (strided), shuffle reduction
      const int warp = tid >> 5, nw = nt >> 5;
      for (int i = warp; i < m; i += nw) {
        const double* row = A + (size_t)(k + 1 + i) * n + (k + 1);
        double acc = 0.0;
        for (int jx = lane; jx < m; jx += 32) acc += row[jx] * v[jx];
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) q[i] = beta * acc;
      }
      __syncthreads();
      double pk = 0.0;
      for (int i = tid; i < m; i += nt) pk += v[i] * q[i];
      const double K = 0.5 * beta * block_sum(pk, sh);
      for (int i = tid; i < m; i += nt) q[i] -= K * v[i];
      __syncthreads();
      for (int idx = tid; idx < m * m; idx += nt) {
        const int i = idx / m, jx = idx % m;
        A[(size_t)(k + 1 + i) * n + (k + 1 + jx)] -= v[i] * q[jx] + q[i] * v[jx];
      }
      for (int i = tid; i < m; i += nt) A[(size_t)(k + 1 + i) * n + k] = v[i];  // keep the reflector
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (n >= 2) { d[n - 2] = A[(size_t)(n - 2) * n + (n - 2)]; e[n - 2] = A[(size_t)(n - 1) * n + (n - 2)]; }
    d[n - 1] = A[(size_t)(n - 1) * n + (n - 1)];
    e[n - 1] = 0.0;
  }
  if (stamps && threadIdx.x == 0) stamps[1] = gtimer();
  // ---- phase 2: Z = Q = H_0 H_1 ... H_{n-3}, accumulated backwards
  for (int idx = tid; idx < n * n; idx += nt) Z[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = n - 3; k >= 0; --k) {
    const double beta = cs[k];
    if (beta == 0.0) continue;  // uniform
    const int m = n - k - 1;
    for (int i = tid; i < m; i += nt) v[i] = A[(size_t)(k + 1 + i) * n + k];
    __syncthreads();
    // w_j = beta * sum_i v_i Z[k+1+i][k+1+j] : column sums; thread per column, coalesced across j
    for (int jx = tid; jx < m; jx += nt) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i) acc += v[i] * Z[(size_t)(k + 1 + i) * n + (k + 1 + jx)];
      q[jx] = beta * acc;
    }
    __syncthreads();
    for (int idx = tid; idx < m * m; idx += nt) {
      const int i = idx / m, jx = idx % m;
      Z[(size_t)(k + 1 + i) * n + (k + 1 + jx)] -= v[i] * q[jx];
    }
    __syncthreads();
  }
  // transpose into V (row i of V = column i of Q) so that plane rotations touch two contiguous rows
  for (int idx = tid; idx < n * n; idx += nt) V[(size_t)(idx % n) * n + (idx / n)] = Z[idx];
  __syncthreads();
  double* Zt = V;

  if (stamps && threadIdx.x == 0) stamps[2] = gtimer();
  // ---- phase 3: implicit QL on (d, e), rotations applied to rows of Zt (tql2)
  const double eps = 2.220446049250313e-16;
  if (tid == 0) { s_scal[0] = 0.0 /* f */; s_scal[1] = 0.0 /* tst1 */; }
  __syncthreads();
  for (int l = 0; l < n; ++l) {
    for (int iter = 0; iter < 64; ++iter) {
      if (tid == 0) {
        if (iter == 0) s_scal[1] = fmax(s_scal[1], fabs(d[l]) + fabs(e[l]));
        int m = l;
        while (m < n - 1 && fabs(e[m]) > eps * s_scal[1]) ++m;
        s_int[0] = m;
        if (m > l) {
          const double g = d[l];
          double p = (d[l + 1] - g) / (2.0 * e[l]);
          double r = hypot(p, 1.0);
          if (p < 0) r = -r;
          d[l] = e[l] / (p + r);
          d[l + 1] = e[l] * (p + r);
          s_scal[2] = g - d[l];  // h
          s_scal[3] = d[l + 1];  // dl1
          s_scal[0] += s_scal[2];
        }
      }
      __syncthreads();
      const int m = s_int[0];
      if (m == l) break;  // uniform
      const double h = s_scal[2];
      for (int i = l + 2 + tid; i < n; i += nt) d[i] -= h;
      __syncthreads();
      if (tid == 0) {
        const double dl1 = s_scal[3], el1 = e[l + 1];
        double p = d[m], c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          const double g = c * e[i], hh = c * p;
          const double r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r; c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = hh + s * (c * g + s * d[i]);
          cs[i] = c; sn[i] = s;
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      }
      __syncthreads();
      // apply the rotations i = m-1 .. l to rows (i, i+1) of Zt; thread k owns column k, carrying row i in a register
      for (int k = tid; k < n; k += nt) {
        double hi = Zt[(size_t)m * n + k];  // current content of row i+1 (starts at row m)
        for (int i = m - 1; i >= l; --i) {
          const double c = cs[i], s = sn[i];
          const double lo = Zt[(size_t)i * n + k];
          Zt[(size_t)(i + 1) * n + k] = s * lo + c * hi;
          hi = c * lo - s * hi;
        }
        Zt[(size_t)l * n + k] = hi;
      }
      __syncthreads();
      if (fabs(e[l]) <= eps * s_scal[1]) break;  // uniform (shared values, read after the barrier)
    }
    __syncthreads();
    if (tid == 0) { d[l] += s_scal[0]; e[l] = 0.0; }
    __syncthreads();
  }
  if (stamps && threadIdx.x == 0) stamps[3] = gtimer();
  // ---- ascending order by rank (ties by index), eigenvectors to the COLUMNS of V via Z as scratch
  for (int i = tid; i < n; i += nt) {
    int r = 0;
    const double di = d[i];
    for (int jx = 0; jx < n; ++jx) r += (d[jx] < di || (d[jx] == di && jx < i)) ? 1 : 0;
    rank_of[i] = r;
    evals[r] = di;
  }
  for (int idx = tid; idx < n * n; idx += nt) Z[idx] = Zt[idx];
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, k = idx % n;  // Z[i][k] = component k of eigenvector i
    V[(size_t)k * n + rank_of[i]] = Z[idx];
  }
}

__global__ void symdecor_scale_kernel(const double* U, const double* evals, int n, double* scaled, int* status) {
  __shared__ double sh[33];
  // min eigenvalue (math.rs:21): block_max pads idle warps with 0, so reduce max(0, 1e-10 - ev) instead of -ev
  double defect = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) defect = fmax(defect, (evals[i] < 1e-10 || evals[i] != evals[i]) ? 1.0 : 0.0);
  const double mn = block_max(defect, sh) > 0.0 ? 0.0 : 1.0;
  if (threadIdx.x == 0) *status = (mn < 1e-10) ? PICARD_SINGULAR_MATRIX : PICARD_OK;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) scaled[e] = U[e] * (1.0 / sqrt(evals[e % n]));
}

inline int ew_blocks(int64_t count) { int64_t b = (count + 255) / 256; return (int)(b > 1184 ? 1184 : (b < 1 ? 1 : b)); }

}  // namespace

#define LAUNCH_CHECK() PICARD_CUDA(cudaGetLastError())

int matmul(const double* A, const double* B, double* C, int n, bool trans_b, double alpha, bool add_identity, cudaStream_t st) {
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  if (trans_b) matmul_kernel<true><<<grid, 256, 0, st>>>(A, B, C, n, n, n, alpha, add_identity ? 1 : 0);
  else matmul_kernel<false><<<grid, 256, 0, st>>>(A, B, C, n, n, n, alpha, add_identity ? 1 : 0);
  LAUNCH_CHECK();
  return 1;
}
int set_identity(double* A, int n, cudaStream_t st) {
  identity_kernel<<<ew_blocks((int64_t)n * n), 256, 0, st>>>(A, n);
  LAUNCH_CHECK();
  return 1;
}
int eye_plus_scaled(const double* D, double alpha, double* M, int n, cudaStream_t st) {
  eye_plus_kernel<<<ew_blocks((int64_t)n * n), 256, 0, st>>>(D, alpha, M, n);
  LAUNCH_CHECK();
  return 1;
}
int copy_scaled(const double* A, double* B, int64_t count, double alpha, cudaStream_t st) {
  scale_kernel<<<ew_blocks(count), 256, 0, st>>>(A, B, count, alpha);
  LAUNCH_CHECK();
  return 1;
}
int iteration_front(const FrontArgs& a, cudaStream_t st) {
  const int nn = a.d.n * a.d.n;
  // the cluster version: n <= 128 (16 CTAs x 8 rows x 128 columns), history of at most 7 pairs in registers, 16-CTA clusters
  // schedulable on this device; PICARD_FRONT_V1=1 keeps the single-CTA kernel (A/B)
  static PerDeviceInt cluster_ok;
  const int use_cluster = cluster_ok.get([&] {  // 1 = yes, 2 = no (0 means "not determined yet" to PerDeviceInt)
    if (getenv("PICARD_FRONT_V1") != nullptr) return 2;
    if (cudaFuncSetAttribute(front_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 2; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(FC_CTAS); cfg.blockDim = dim3(FC_THREADS); cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = FC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, front_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 2; }
    return nclusters >= 1 ? 1 : 2;
  });
  if (use_cluster == 1 && a.d.n <= 128 && a.d.m <= FC_M) {
    front_cluster_kernel<<<FC_CTAS, FC_THREADS, 0, st>>>(a);
    LAUNCH_CHECK();
    return 1;
  }
  int threads = nn >= 1024 ? 1024 : ((nn + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  front_kernel<<<1, threads, 0, st>>>(a);
  LAUNCH_CHECK();
  return 1;
}
int loss_from_moments(const CoreDims& d, const double* mom, const double* signs, CoreScalars* sc, int which, cudaStream_t st,
                      CoreScalars* sc_map, unsigned long long seq) {
  loss_kernel<<<1, 32, 0, st>>>(d, mom, signs, sc, which, sc_map, seq);
  LAUNCH_CHECK();
  return 1;
}
int accept_step(const CoreDims& d, const double* D, double alpha, double* S_prev, const double* W, double* C, int update_c,
                CoreScalars* sc, cudaStream_t st) {
  const int64_t nn = (int64_t)d.n * d.n;
  accept_kernel<<<ew_blocks(nn), 256, 0, st>>>(D, alpha, S_prev, nn, sc);
  LAUNCH_CHECK();
  int launches = 1;
  if (update_c) launches += matmul(W, W, C, d.n, true, 1.0, false, st);
  return launches;
}
int negate_into(const double* G, double* D, int64_t count, CoreScalars* sc, cudaStream_t st) {
  negate_kernel<<<ew_blocks(count), 256, 0, st>>>(G, D, count, sc);
  LAUNCH_CHECK();
  return 1;
}

int matrix_exp(const double* D, double alpha, double norm_d, int n, const ExpmWork& w, double* out, cudaStream_t st, const double* W,
               double* Wt) {
  const double norm = norm_d * alpha;  // max |D * alpha| (alpha > 0)
  if (!(norm >= 1e-15)) {  // math.rs:43-45 (NaN norms cannot occur: fmax ignores NaN entries)
    int launches = 0;
    if (out) launches += set_identity(out, n, st);
    if (W && Wt) { PICARD_CUDA(cudaMemcpyAsync(Wt, W, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, st)); }
    return launches;
  }
  int s = (int)std::fmax(std::ceil(std::log2(norm)), 0.0);
  double inv_scale = std::ldexp(1.0, -s);  // dividing by 2^s == multiplying by 2^-s, exactly
  double first_norm = norm * inv_scale;
  const int grid = (n + EXPM_ROWS - 1) / EXPM_ROWS;
  if (grid > EXPM_MAX_CTAS) throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: matrix_exp supports n <= 256");
  int as_in_smem = n <= 128 ? 1 : 0;  // A_s (n x n) staged in shared memory when it fits
  const size_t smem = sizeof(double) * (3 * EXPM_ROWS * (size_t)(n + 4) + (as_in_smem ? (size_t)n * (n + 8) : 0));
  static PerDeviceInt configured;
  configured.get([&] {
    PICARD_CUDA(cudaFuncSetAttribute(expm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(double) * (3 * EXPM_ROWS * (128 + 4) + 128 * (128 + 8)))));
    return 1;
  });
  // w.slots: [barrier counter (8 doubles)][flags: 31 x grid doubles]; all-ones bytes = NaN = "not published yet"
  unsigned int* bar = reinterpret_cast<unsigned int*>(w.slots);
  double* flags = w.slots + 8;
  PICARD_CUDA(cudaMemsetAsync(bar, 0, sizeof(double) * 8, st));
  PICARD_CUDA(cudaMemsetAsync(flags, 0xFF, sizeof(double) * 31 * (size_t)grid, st));
  double* r0 = w.res0; double* r1 = w.res1;
  void* args[] = {(void*)&D, (void*)&alpha, (void*)&inv_scale, (void*)&first_norm, (void*)&s, (void*)&n, (void*)&flags, (void*)&bar,
                  (void*)&r0, (void*)&r1, (void*)&out, (void*)&W, (void*)&Wt, (void*)&as_in_smem};
  PICARD_CUDA(cudaLaunchCooperativeKernel((const void*)expm_rows_kernel, dim3(grid), dim3(256), args, smem, st));
  return 1;
}

int matrix_exp_candidates(const double* D, double alpha0, double norm_d, int n, int n_cand, const ExpmWork& w, const double* W, double* Wt_all,
                          cudaStream_t st) {
  const double norm = norm_d * alpha0;
  if (!(norm >= 1e-15) || norm > 1.0 || n_cand < 2 || n > 256) return -1;  // s > 0 or degenerate: the caller uses the per-try path
  if (n_cand > EXPM_NC) n_cand = EXPM_NC;
  double first_norm = norm;
  const int grid = (n + EXPM_ROWS - 1) / EXPM_ROWS;
  int as_in_smem = n <= 128 ? 1 : 0;
  const size_t smem = sizeof(double) * (2 * EXPM_ROWS * (size_t)(n + 4) + (as_in_smem ? (size_t)n * (n + 8) : 0));
  static PerDeviceInt configured;
  configured.get([&] {
    const int mx = (int)(sizeof(double) * (2 * EXPM_ROWS * (128 + 4) + 128 * (128 + 8)));
    PICARD_CUDA(cudaFuncSetAttribute(expm_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    PICARD_CUDA(cudaFuncSetAttribute(expm_multi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    return 1;
  });
  double* flags = w.slots + 8;
  PICARD_CUDA(cudaMemsetAsync(flags, 0xFF, sizeof(double) * 31 * (size_t)grid, st));
  void* args[] = {(void*)&D, (void*)&alpha0, (void*)&first_norm, (void*)&n, (void*)&n_cand, (void*)&flags, (void*)&W, (void*)&Wt_all,
                  (void*)&as_in_smem};
  if (n <= 128) PICARD_CUDA(cudaLaunchCooperativeKernel((const void*)expm_multi_kernel<2>, dim3(grid), dim3(256), args, smem, st));
  else PICARD_CUDA(cudaLaunchCooperativeKernel((const void*)expm_multi_kernel<4>, dim3(grid), dim3(256), args, smem, st));
  return n_cand;
}

int sln_det(const double* A, int n, double* work, double* out2, cudaStream_t st) {
  int threads = n >= 48 ? 1024 : 256;
  lu_logdet_kernel<<<1, threads, 0, st>>>(A, n, work, out2);
  LAUNCH_CHECK();
  return 1;
}

int jacobi_eigh(double* A, int n, double* V, double* evals, double* scratch, cudaStream_t st) {
  if (n > EIG_MAX_N) throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: eigendecomposition supports n <= 512");
  if (n > 32) {  // Householder + implicit QL: O(n^3) data-parallel work, only the O(n^2) QL recurrence is sequential
    double* z = scratch;  // caller-provided: n^2 + 8 doubles (no allocation here: driver allocator calls after an idle period
                          // were measured at up to ~1 s on the B200 hosts, 100x the kernel itself)
    unsigned long long* stamps = getenv("PICARD_TRACE") ? reinterpret_cast<unsigned long long*>(z + (size_t)n * n) : nullptr;
    tridiag_ql_kernel<<<1, BT1, 0, st>>>(A, n, z, V, evals, stamps);
    LAUNCH_CHECK();
    if (stamps) {
      unsigned long long h[4];
      PICARD_CUDA(cudaMemcpyAsync(h, stamps, sizeof h, cudaMemcpyDeviceToHost, st));
      PICARD_CUDA(cudaStreamSynchronize(st));
      fprintf(stderr, "[picard trace]     eigh n=%d: tridiag %.3f ms, accumulate Q %.3f ms, QL %.3f ms\n", n, (h[1] - h[0]) * 1e-6,
              (h[2] - h[1]) * 1e-6, (h[3] - h[2]) * 1e-6);
    }
    return 1;
  }
  int threads = n * n >= 1024 ? 1024 : ((n * n + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  jacobi_kernel<<<1, threads, 0, st>>>(A, n, V, evals, scratch);
  LAUNCH_CHECK();
  return 1;
}

namespace {
__global__ void fastica_a_kernel(const double* __restrict__ mom, int n, double t_total, double* __restrict__ A) {
  const double* Gr = mom + mom_off_gr(n);
  const double* Sd = mom + mom_off_sd(n);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int i = e / n, j = e % n;
    double v = Gr[e] / t_total;
    if (i == j) v -= Sd[i] / t_total;
    A[e] = v;
  }
}
}  // namespace

int fastica_matrix(const double* mom, int n, double t_total, const double* W, double* tmp, double* C, cudaStream_t st) {
  fastica_a_kernel<<<ew_blocks((int64_t)n * n), 256, 0, st>>>(mom, n, t_total, tmp);
  LAUNCH_CHECK();
  return 1 + matmul(tmp, W, C, n, false, 1.0, false, st);
}

int sym_decorrelation(const double* W, int n, double* work, double* out, int* status_dev, cudaStream_t st) {
  double* wwt = work;
  double* U = work + (size_t)n * n;
  double* scaled = work + 2 * (size_t)n * n;
  double* t2 = work + 3 * (size_t)n * n;
  double* ev = work + 4 * (size_t)n * n;
  double* scratch = ev + n + (n & 1);  // n^2 + 8 doubles for the eigensolver
  int launches = matmul(W, W, wwt, n, true, 1.0, false, st);
  launches += jacobi_eigh(wwt, n, U, ev, scratch, st);
  int threads = n * n >= 1024 ? 1024 : ((n * n + 31) / 32) * 32;
  symdecor_scale_kernel<<<1, threads, 0, st>>>(U, ev, n, scaled, status_dev);
  LAUNCH_CHECK();
  ++launches;
  launches += matmul(scaled, U, t2, n, true, 1.0, false, st);
  launches += matmul(t2, W, out, n, false, 1.0, false, st);
  return launches;
}

}  // namespace small
}  // namespace picard
