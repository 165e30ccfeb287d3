// step_instances.cu -- explicit instantiation of the step-kernel variants of one (KFUSED, P1BOT, P2BOT) combination
// (x dense / sparse reward x step mask x two CTA shapes = 8 kernels).  Compiled eight times by build.py with
// -DFG_INST_KF=.. -DFG_INST_B1=.. -DFG_INST_B2=.. so that the kernels compile in parallel.
#define FG_STEP_INSTANCE
#include "step_kernel.cuh"

namespace fgk {
template cudaError_t launch_step_d<(FG_INST_KF != 0), (FG_INST_B1 != 0), (FG_INST_B2 != 0)>(bool, bool, int, cudaStream_t, const Params &);
}
