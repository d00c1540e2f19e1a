#include "rowblock_inst.cuh"
namespace picard { template int launch_rb_loss<64>(const PassLaunch&, const CUtensorMap&); }
