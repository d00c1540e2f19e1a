// Host-side cost of the CUDA API calls a line-search try makes, on an idle stream (each call followed by a sync so the
// launch latency is exposed, as in the solver's host-driven loop), vs one graph launch of the same five nodes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/api_latency profiles/microbench/api_latency.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tiny(double* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  double* d; cudaMalloc(&d, 1024); cudaMemset(d, 0, 1024);
  double* h; cudaMallocHost(&h, 64);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int R = 2000;
  auto bench = [&](const char* name, auto f) {
    f(); cudaStreamSynchronize(st);
    double t0 = now_us();
    for (int i = 0; i < R; ++i) { f(); cudaStreamSynchronize(st); }
    printf("{\"op\": \"%s\", \"us_per_iter_incl_sync\": %.2f}\n", name, (now_us() - t0) / R);
  };
  bench("sync only", [&] {});
  bench("memsetAsync 64B", [&] { cudaMemsetAsync(d, 0, 64, st); });
  bench("kernel launch", [&] { tiny<<<1, 32, 0, st>>>(d); });
  bench("cooperative launch (16 CTAs)", [&] { void* a[] = {(void*)&d}; cudaLaunchCooperativeKernel((const void*)tiny, dim3(16), dim3(32), a, 0, st); });
  bench("memcpyAsync D2H 64B", [&] { cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, st); });
  bench("try-like: 2 memset + coop + 3 kernels + 2 events + D2H", [&] {
    static cudaEvent_t e0 = nullptr, e1 = nullptr; if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    cudaMemsetAsync(d, 0, 64, st); cudaMemsetAsync(d + 8, 0, 64, st);
    void* a[] = {(void*)&d}; cudaLaunchCooperativeKernel((const void*)tiny, dim3(16), dim3(32), a, 0, st);
    cudaEventRecord(e0, st); tiny<<<148, 256, 0, st>>>(d); tiny<<<4, 256, 0, st>>>(d); cudaEventRecord(e1, st); tiny<<<1, 32, 0, st>>>(d);
    cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, st);
  });
  // the same as a graph
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  { cudaMemsetAsync(d, 0, 64, st); cudaMemsetAsync(d + 8, 0, 64, st);
    void* a[] = {(void*)&d}; cudaError_t e = cudaLaunchCooperativeKernel((const void*)tiny, dim3(16), dim3(32), a, 0, st);
    if (e != cudaSuccess) printf("{\"note\": \"cooperative launch not capturable: %s\"}\n", cudaGetErrorString(e));
    tiny<<<148, 256, 0, st>>>(d); tiny<<<4, 256, 0, st>>>(d); tiny<<<1, 32, 0, st>>>(d);
    cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, st); }
  cudaError_t ce = cudaStreamEndCapture(st, &g);
  if (ce == cudaSuccess && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) bench("graph launch of the try-like sequence", [&] { cudaGraphLaunch(ge, st); });
  else printf("{\"note\": \"graph capture failed: %s\"}\n", cudaGetErrorString(ce));
  // mapped pinned polling instead of memcpy + sync
  volatile double* hm; cudaHostAlloc((void**)&hm, 64, cudaHostAllocMapped); double* dm; cudaHostGetDevicePointer((void**)&dm, (void*)hm, 0);
  hm[0] = 0;
  { double t0 = now_us(); double expect = 0;
    for (int i = 0; i < R; ++i) { expect += 1.0; tiny<<<1, 32, 0, st>>>(dm); while (hm[0] != expect) {} }
    printf("{\"op\": \"kernel launch + poll mapped host memory\", \"us_per_iter_incl_sync\": %.2f}\n", (now_us() - t0) / R); }
  return 0;
}
