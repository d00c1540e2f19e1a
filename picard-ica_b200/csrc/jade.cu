#include "jade.cuh"
namespace picard {
void jade_device(const double*, int, int64_t, int64_t, double, int64_t, double, bool, picard_comm*, int, cudaStream_t, double*, int64_t*,
                 picard_stats_t*) {
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the JADE warm start is not implemented on the device yet");
}
void jade_cumulants_device(const double*, int, int64_t, int64_t, double, picard_comm*, int, cudaStream_t, double*, picard_stats_t*) {
  throw Error(PICARD_COMPUTATION_ERROR, "Computation error: the JADE warm start is not implemented on the device yet");
}
}  // namespace picard
