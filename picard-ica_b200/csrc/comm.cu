// Sample-axis communicator: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// NCCL is loaded lazily with dlopen so single-GPU users carry no NCCL dependency; inside a Python process
// that already imported torch this resolves to torch's bundled libnccl.so.2, otherwise to the system one.
#include "engine.cuh"
#include "p2p.cuh"

#include <dlfcn.h>

#include <mutex>
#include <vector>

namespace picard {

namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GroupFn)(void);

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  AllGatherFn all_gather = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (a.handle) {
      a.get_unique_id = (GetUniqueIdFn)dlsym(a.handle, "ncclGetUniqueId");
      a.comm_init_rank = (CommInitRankFn)dlsym(a.handle, "ncclCommInitRank");
      a.comm_destroy = (CommDestroyFn)dlsym(a.handle, "ncclCommDestroy");
      a.all_reduce = (AllReduceFn)dlsym(a.handle, "ncclAllReduce");
      a.all_gather = (AllGatherFn)dlsym(a.handle, "ncclAllGather");
      a.get_error_string = (GetErrorStringFn)dlsym(a.handle, "ncclGetErrorString");
      a.group_start = (GroupFn)dlsym(a.handle, "ncclGroupStart");
      a.group_end = (GroupFn)dlsym(a.handle, "ncclGroupEnd");
    }
  });
  if (!a.handle || !a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce)
    throw Error(PICARD_COMPUTATION_ERROR, "Computation error: NCCL (libnccl.so.2) could not be loaded");
  return a;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    NcclApi& a = api();
    throw Error(PICARD_COMPUTATION_ERROR, std::string("Computation error: NCCL failure in ") + what + ": " +
                                              (a.get_error_string ? a.get_error_string(rc) : "unknown"));
  }
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
}  // namespace

void comm_unique_id(char id[PICARD_UNIQUE_ID_BYTES]) {
  NcclUniqueId u;
  nccl_check(api().get_unique_id(&u), "ncclGetUniqueId");
  memcpy(id, u.internal, 128);
}

}  // namespace picard

struct picard_comm {
  void* nccl = nullptr;
  int rank = 0, nranks = 1, device = 0;
  bool p2p = false;
  void* p2p_local = nullptr;         // this rank's mailbox allocation (flags first, then the slots)
  picard::P2PPeers peers{};
  unsigned long long p2p_calls = 0;  // allreduces issued so far (parity = calls & 1)
};

namespace picard {

namespace {
constexpr size_t P2P_FLAG_BYTES = 4096;  // [2][P2P_MAX_RANKS] counters, padded
constexpr size_t P2P_BOX_BYTES = P2P_FLAG_BYTES + sizeof(double) * 2 * P2P_MAX_RANKS * P2P_MAX_DOUBLES;

// One-shot allreduce (sum, f64) of a small buffer over NVLink peer memory, ONE launch per call, bit-identical on every rank:
//   push : CTA (dst, part) copies part `part` of the local buffer into rank dst's mailbox slot [parity][this rank] (16-byte peer
//          stores), fences at system scope and bumps dst's arrival counter [parity][this rank];
//   sum  : once all P2P_PARTS parts of every source have arrived in the local mailbox, every CTA sums its slice over the source
//          ranks IN RANK ORDER (the same order on every rank: the replicated N x N state stays bit-identical) and writes it back.
// Counters only grow (call k on a parity expects k * P2P_PARTS); the two parities alternate, so a rank that is already pushing
// call i + 1 never overwrites slots a slower rank is still summing for call i (it cannot reach call i + 2 before that rank
// has pushed call i + 1, i.e. finished call i).
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(double* __restrict__ buf, int count, int rank, int nranks, P2PPeers peers, int parity,
                                                            unsigned int expect) {
  const int dst = blockIdx.x % nranks, part = blockIdx.x / nranks;
  const int n2 = (count + 1) / 2;                                   // double2 units (the buffers are 16-byte aligned, padded)
  const int per = (n2 + P2P_PARTS - 1) / P2P_PARTS;
  const int lo = part * per, hi = lo + per < n2 ? lo + per : n2;
  {
    double2* out = reinterpret_cast<double2*>(peers.box[dst] + p2p_slot(parity, rank));
    const double2* in = reinterpret_cast<const double2*>(buf);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = in[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd_system(peers.flags[dst] + parity * P2P_MAX_RANKS + rank, 1u);
  }
  // wait for every source's parts in the LOCAL mailbox
  if (threadIdx.x < nranks) {
    volatile unsigned int* f = peers.flags[rank] + parity * P2P_MAX_RANKS + threadIdx.x;
    while ((int)(*f - expect) < 0) __nanosleep(32);
    __threadfence_system();
  }
  __syncthreads();
  // this CTA's slice of the result: gridDim.x slices of double2 units
  const int nsl = gridDim.x, sper = (n2 + nsl - 1) / nsl;
  const int slo = blockIdx.x * sper, shi = slo + sper < n2 ? slo + sper : n2;
  const double2* box = reinterpret_cast<const double2*>(peers.box[rank] + p2p_slot(parity, 0));
  const size_t stride2 = P2P_MAX_DOUBLES / 2;
  for (int i = slo + threadIdx.x; i < shi; i += blockDim.x) {
    double2 acc = __ldcv(box + i);
    for (int q = 1; q < nranks; ++q) { const double2 v = __ldcv(box + (size_t)q * stride2 + i); acc.x += v.x; acc.y += v.y; }
    reinterpret_cast<double2*>(buf)[i] = acc;
  }
}

// Maps every peer's mailbox (CUDA IPC handles exchanged with ncclAllGather).  Any failure leaves c->p2p false: NCCL is used.
void p2p_setup(picard_comm* c) {
  if (c->nranks < 2 || c->nranks > P2P_MAX_RANKS || getenv("PICARD_NO_P2P") != nullptr || !api().all_gather) return;
  static_assert(P2P_MAX_DOUBLES % 2 == 0, "slots must keep 16-byte alignment");
  void* local = nullptr;
  cudaIpcMemHandle_t* d_handles = nullptr;
  bool ok = cudaMalloc(&local, P2P_BOX_BYTES) == cudaSuccess && cudaMemset(local, 0, P2P_BOX_BYTES) == cudaSuccess &&
            cudaMalloc((void**)&d_handles, sizeof(cudaIpcMemHandle_t) * c->nranks) == cudaSuccess;
  std::vector<cudaIpcMemHandle_t> handles((size_t)c->nranks);
  if (ok) {
    cudaIpcMemHandle_t mine;
    ok = cudaIpcGetMemHandle(&mine, local) == cudaSuccess &&
         cudaMemcpy(d_handles + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  // collective: every rank takes part even if its own setup failed (a zero handle then makes the peers fall back too)
  int rc = 1;
  if (d_handles) {
    rc = api().all_gather(d_handles + c->rank, d_handles, sizeof(cudaIpcMemHandle_t), /*ncclChar*/ 0, c->nccl, 0);
    if (rc == 0 && cudaMemcpy(handles.data(), d_handles, sizeof(cudaIpcMemHandle_t) * c->nranks, cudaMemcpyDeviceToHost) != cudaSuccess) rc = 1;
  }
  ok = ok && rc == 0;
  if (ok) {
    for (int q = 0; q < c->nranks && ok; ++q) {
      void* base = local;
      if (q != c->rank) ok = cudaIpcOpenMemHandle(&base, handles[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      c->peers.flags[q] = reinterpret_cast<unsigned int*>(base);
      c->peers.box[q] = reinterpret_cast<double*>(static_cast<char*>(base) + P2P_FLAG_BYTES);
    }
  }
  if (d_handles) cudaFree(d_handles);
  cudaGetLastError();
  // agree: one NCCL allreduce of the success flags (a rank that failed must switch every rank back to NCCL)
  double* d_ok = nullptr;
  double h_ok = ok ? 1.0 : 0.0;
  if (cudaMalloc((void**)&d_ok, sizeof(double)) == cudaSuccess) {
    cudaMemcpy(d_ok, &h_ok, sizeof(double), cudaMemcpyHostToDevice);
    if (api().all_reduce(d_ok, d_ok, 1, kNcclFloat64, kNcclSum, c->nccl, 0) != 0) h_ok = 0.0;
    else cudaMemcpy(&h_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_ok);
  } else h_ok = 0.0;
  cudaGetLastError();
  c->p2p_local = local;
  c->p2p = (h_ok == (double)c->nranks);
  if (getenv("PICARD_TRACE") != nullptr && c->rank == 0)
    fprintf(stderr, "[picard trace] peer-memory allreduce: %s\n", c->p2p ? "enabled (CUDA IPC mailboxes)" : "unavailable, using NCCL");
}
}  // namespace

picard_comm* comm_create(const char id[PICARD_UNIQUE_ID_BYTES], int rank, int nranks, int device) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(PICARD_INVALID_CONFIG, "Invalid configuration for 'comm': bad rank / size");
  PICARD_CUDA(cudaSetDevice(device));
  NcclUniqueId u;
  memcpy(u.internal, id, 128);
  picard_comm* c = new picard_comm();
  c->rank = rank; c->nranks = nranks; c->device = device;
  int rc = api().comm_init_rank(&c->nccl, nranks, u, rank);
  if (rc != 0) { delete c; nccl_check(rc, "ncclCommInitRank"); }
  // first collective on a fresh communicator sets up channels / NVLS buffers (tens of ms): do it here, not inside the first fit
  if (nranks > 1) {
    double* tmp = nullptr;
    if (cudaMalloc(&tmp, sizeof(double) * 1024) == cudaSuccess) {
      cudaMemset(tmp, 0, sizeof(double) * 1024);
      for (size_t cnt : {(size_t)1, (size_t)256, (size_t)1024})
        api().all_reduce(tmp, tmp, cnt, kNcclFloat64, kNcclSum, c->nccl, 0);
      cudaStreamSynchronize(0);
      cudaFree(tmp);
    }
    p2p_setup(c);
  }
  return c;
}
void comm_destroy(picard_comm* c) {
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->p2p_local) {
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank && c->peers.flags[q] != nullptr) cudaIpcCloseMemHandle(c->peers.flags[q]);
    if (c->nccl && c->nranks > 1) {  // nobody frees a mailbox a peer may still be writing to
      double* tmp = nullptr;
      if (cudaMalloc((void**)&tmp, sizeof(double)) == cudaSuccess) {
        cudaMemset(tmp, 0, sizeof(double));
        api().all_reduce(tmp, tmp, 1, kNcclFloat64, kNcclSum, c->nccl, 0);
        cudaStreamSynchronize(0);
        cudaFree(tmp);
      }
    }
    cudaFree(c->p2p_local);
  }
  cudaGetLastError();
  if (c->nccl) api().comm_destroy(c->nccl);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
}
int comm_rank(const picard_comm* c) { return c ? c->rank : 0; }
int comm_size(const picard_comm* c) { return c ? c->nranks : 1; }

void comm_allreduce_sum(picard_comm* c, double* d_buf, size_t count, cudaStream_t st) {
  if (!c || c->nranks == 1 || count == 0) return;
  if (c->p2p && count <= P2P_MAX_DOUBLES - 2 && (reinterpret_cast<uintptr_t>(d_buf) & 15) == 0) {
    const int parity = (int)(c->p2p_calls & 1);
    const unsigned int expect = (unsigned int)((c->p2p_calls / 2 + 1) * P2P_PARTS);
    ++c->p2p_calls;
    p2p_allreduce_kernel<<<c->nranks * P2P_PARTS, 512, 0, st>>>(d_buf, (int)count, c->rank, c->nranks, c->peers, parity, expect);
    PICARD_CUDA(cudaGetLastError());
    return;
  }
  nccl_check(api().all_reduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
}
// For kernels that carry the exchange in their own tail: the next call's parity / expected counter value (advances the call count).
bool comm_p2p_next(picard_comm* c, size_t count, P2PCall* out) {
  if (!c || c->nranks == 1 || !c->p2p || count > P2P_MAX_DOUBLES - 2) return false;
  out->peers = c->peers; out->rank = c->rank; out->nranks = c->nranks;
  out->parity = (int)(c->p2p_calls & 1);
  out->expect = (unsigned int)((c->p2p_calls / 2 + 1) * P2P_PARTS);
  ++c->p2p_calls;
  return true;
}

void comm_allreduce_sum2(picard_comm* c, double* a, size_t na, double* b, size_t nb, cudaStream_t st) {
  if (!c || c->nranks == 1) return;
  if (c->p2p) {  // two one-shot exchanges (each one launch)
    if (na) comm_allreduce_sum(c, a, na, st);
    if (nb) comm_allreduce_sum(c, b, nb, st);
    return;
  }
  NcclApi& A = api();
  if (A.group_start && A.group_end) nccl_check(A.group_start(), "ncclGroupStart");
  if (na) nccl_check(A.all_reduce(a, a, na, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
  if (nb) nccl_check(A.all_reduce(b, b, nb, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
  if (A.group_start && A.group_end) nccl_check(A.group_end(), "ncclGroupEnd");
}

}  // namespace picard
