// agx_step_balloon.cu — instantiates the fused step kernel (agx_step_kernel.cuh) for the balloon task, all control modes.
#include "agx_step_kernel.cuh"

namespace agxk {
template int agx_dispatch_task<AGX_TASK_BALLOON>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
template int agx_observe_task<AGX_TASK_BALLOON>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
}  // namespace agxk
