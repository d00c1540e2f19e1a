#include "rowblock_inst.cuh"
namespace picard { template int launch_rb_loss<128>(const PassLaunch&, const CUtensorMap&); }

#ifdef PICARD_RB_TRACE
extern "C" int picard_debug_rb_trace(long long* out, int n) {
  const int total = 16 * picard::RB_TRACE_TILES * 4;
  if (n > total) n = total;
  return (int)cudaMemcpyFromSymbol(out, picard::g_rb_trace, (size_t)n * sizeof(long long));
}
#endif
