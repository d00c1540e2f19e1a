#include "rowblock_inst.cuh"
namespace picard { template int launch_rb_grady<16>(const PassLaunch&, const CUtensorMap&); }
