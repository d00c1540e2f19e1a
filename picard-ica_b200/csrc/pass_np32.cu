#include "pass_inst.cuh"
namespace picard { template int launch_pass_np<32>(const PassLaunch&, const CUtensorMap&); }
