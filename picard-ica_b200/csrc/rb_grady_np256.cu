#include "rowblock_inst.cuh"
namespace picard { template int launch_rb_grady<256>(const PassLaunch&, const CUtensorMap&); }
