// Sample-axis communicator: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// NCCL is loaded lazily with dlopen so single-GPU users carry no NCCL dependency; inside a Python process
// that already imported torch this resolves to torch's bundled libnccl.so.2, otherwise to the system one.
#include "engine.cuh"

#include <dlfcn.h>

#include <mutex>

namespace picard {

namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GroupFn)(void);

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
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
};

namespace picard {

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
  }
  return c;
}
void comm_destroy(picard_comm* c) {
  if (!c) return;
  if (c->nccl) api().comm_destroy(c->nccl);
  delete c;
}
int comm_rank(const picard_comm* c) { return c ? c->rank : 0; }
int comm_size(const picard_comm* c) { return c ? c->nranks : 1; }

void comm_allreduce_sum(picard_comm* c, double* d_buf, size_t count, cudaStream_t st) {
  if (!c || c->nranks == 1 || count == 0) return;
  nccl_check(api().all_reduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
}
void comm_allreduce_sum2(picard_comm* c, double* a, size_t na, double* b, size_t nb, cudaStream_t st) {
  if (!c || c->nranks == 1) return;
  NcclApi& A = api();
  if (A.group_start && A.group_end) nccl_check(A.group_start(), "ncclGroupStart");
  if (na) nccl_check(A.all_reduce(a, a, na, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
  if (nb) nccl_check(A.all_reduce(b, b, nb, kNcclFloat64, kNcclSum, c->nccl, st), "ncclAllReduce");
  if (A.group_start && A.group_end) nccl_check(A.group_end(), "ncclGroupEnd");
}

}  // namespace picard
