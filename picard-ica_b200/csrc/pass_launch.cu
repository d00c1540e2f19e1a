// Host side of the fused pass: TMA tensor map over the sample matrix and dispatch on the padded size.
#include <mutex>

#include "pass.cuh"

namespace picard {

int pass_padded_size(int n) {
  if (n <= 0) throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: Input matrix cannot be empty");
  for (int np : {8, 16, 32, 64, 128, 256})
    if (n <= np) return np;
  throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: more than 256 components / features are not supported by this build");
}

size_t pass_workspace_doubles(int n, int sm_count) {
  const int np = pass_padded_size(n);
  // per-CTA partials: the fused / grad kernels write NP x NP (x2 with H) per CTA, the row-block kernels RP x NP
  if (np <= 64) return (size_t)sm_count * PASS_MAX_BLOCKS_PER_SM * (size_t)pass_partial_size(np, true, true);
  if (np == 128) return (size_t)sm_count * 2 * (size_t)pass_partial_size(128, true, true);
  return (size_t)sm_count * 2 * (size_t)(2 * 64 * 256 + 3 * 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  if (!fn) throw Error(PICARD_COMPUTATION_ERROR, "Computation error: cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

// A (n_rows x t_local) f64 matrix, row-major, leading dimension ldx.  Box = box_cols samples x box_rows rows; SWIZZLE_128B
// (box_cols = 16: a 128-byte swizzle row) or dense rows.  Rows >= n_rows and samples >= t_local are zero-filled on loads and
// clipped on stores by the TMA unit.
CUtensorMap make_tmap_box(const double* d_x, int64_t ldx, int64_t t_local, int n_rows, int box_cols, int box_rows, bool swizzle128) {
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0 || (ldx & 1) != 0)
    throw Error(PICARD_INVALID_DIMENSIONS,
                "Invalid dimensions: device sample matrix must be 16-byte aligned with an even row stride (TMA)");
  if (t_local <= 0 || t_local >= (int64_t)1 << 31)
    throw Error(PICARD_INVALID_DIMENSIONS, "Invalid dimensions: per-GPU sample count must be in [1, 2^31)");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)t_local, (cuuint64_t)n_rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ldx * 8};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(d_x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// Box = 16 samples x NP rows, SWIZZLE_128B: what the pass kernels load
CUtensorMap make_tmap(const double* d_x, int64_t ldx, int64_t t_local, int n_in, int np) {
  return make_tmap_box(d_x, ldx, t_local, n_in, 16, np, true);
}

int launch_pass(const PassLaunch& L) {
  const int nmax = L.n_in > L.n_out ? L.n_in : L.n_out;
  const int np = pass_padded_size(nmax);
  CUtensorMap tmap = make_tmap(L.d_x, L.ldx, L.t_local, L.n_in, np);
  // Gram-type passes read their input as "Y": the stored-Y gradient pass, and psi(y) = y (covariance of the whitening step,
  // C = Y Y^T / T of core.rs:202)
  if (L.mode == PASS_GRADY || (L.mode == PASS_GRAD && L.dens == DENS_LINEAR)) {
    switch (np) {
      case 8: return launch_rb_grady<8>(L, tmap);
      case 16: return launch_rb_grady<16>(L, tmap);
      case 32: return launch_rb_grady<32>(L, tmap);
      case 64: return launch_rb_grady<64>(L, tmap);
      case 128: return launch_rb_grady<128>(L, tmap);
      default: return launch_rb_grady<256>(L, tmap);
    }
  }
  if ((L.mode == PASS_LOSS || L.mode == PASS_APPLY) && np >= 64)
    return np == 64 ? launch_rb_loss<64>(L, tmap) : (np == 128 ? launch_rb_loss<128>(L, tmap) : launch_rb_loss<256>(L, tmap));
  if (np > 128)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: N > 128 needs the Y store (the from-X gradient kernels hold N x N accumulators "
                                          "in one SM's registers); do not set PICARD_FLAG_NO_Y_STORE / free some device memory");
  switch (np) {
    case 8: return launch_pass_np<8>(L, tmap);
    case 16: return launch_pass_np<16>(L, tmap);
    case 32: return launch_pass_np<32>(L, tmap);
    case 64: return launch_pass_np<64>(L, tmap);
    default: return launch_pass_np<128>(L, tmap);
  }
}

}  // namespace picard
