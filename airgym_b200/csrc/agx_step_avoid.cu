// agx_step_avoid.cu — instantiates the fused step kernel (agx_step_kernel.cuh) for the avoid task, all control modes.
#include "agx_step_kernel.cuh"

namespace agxk {
template int agx_dispatch_task<AGX_TASK_AVOID>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
template int agx_observe_task<AGX_TASK_AVOID>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
}  // namespace agxk
