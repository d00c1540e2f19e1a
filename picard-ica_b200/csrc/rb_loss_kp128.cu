#include "rowblock_inst.cuh"
namespace picard { template int launch_rb_loss<128>(const PassLaunch&, const CUtensorMap&); }
